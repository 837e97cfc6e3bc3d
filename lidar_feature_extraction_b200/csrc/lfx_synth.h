// lfx_synth.h — deterministic synthetic LiDAR scans (host + device), used by tests and bench.py.
// Not part of the reference; shapes and worlds follow BASELINE.json `configs` / SURVEY.md §8(d).
// One function computes one return so that the host loop and the CUDA kernel share the scene.
#ifndef LFX_SYNTH_H_
#define LFX_SYNTH_H_

#include <math.h>
#include <stdint.h>

#include "lfx.h"

#ifdef __CUDACC__
#define LFX_HD __host__ __device__ __forceinline__
#else
#define LFX_HD inline
#endif

namespace lfx_synth
{

struct Point32  // deployed wire layout, point_type_converter/convert.py:137-145
{
  float x, y, z, w;
  float intensity;
  uint16_t ring;
  uint16_t pad0;
  uint32_t pad1[2];
};
static_assert(sizeof(Point32) == 32, "wire point must be 32 bytes");

LFX_HD uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// counter-based: one 64-bit draw per (seed, frame, ring, col, stream)
LFX_HD uint64_t draw(uint64_t seed, uint64_t frame, uint32_t ring, uint32_t col, uint32_t stream)
{
  uint64_t h = mix64(seed ^ mix64(frame));
  h = mix64(h ^ (((uint64_t)ring << 40) | ((uint64_t)col << 8) | stream));
  return h;
}

LFX_HD float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }  // [0,1)

struct Pose { float px, py, yaw; };

LFX_HD Pose pose_of(const lfx_synth_spec & s, uint64_t frame)
{
  const float f = (float)(frame % 100000ull);
  Pose p;
  if (s.world == LFX_WORLD_TUNNEL) {
    p.px = fmodf(0.1f * f, 40.0f) - 20.0f;
    p.py = 0.5f * sinf(0.05f * f);
    p.yaw = 0.05f * sinf(0.031f * f);
  } else {
    p.px = 3.0f * sinf(0.013f * f);
    p.py = 1.5f * sinf(0.007f * f + 1.0f);
    p.yaw = 0.2f * sinf(0.011f * f);
  }
  return p;
}

LFX_HD float slab_exit(float o, float d, float lo, float hi)
{
  if (d > 0.0f) { return (hi - o) / d; }
  if (d < 0.0f) { return (lo - o) / d; }
  return 1e30f;
}

LFX_HD float hit_cylinder(float ox, float oy, float dx, float dy, float cx, float cy, float rad)
{
  const float a = dx * dx + dy * dy;
  if (a < 1e-12f) { return 1e30f; }
  const float fx = ox - cx, fy = oy - cy;
  const float b = fx * dx + fy * dy;
  const float c = fx * fx + fy * fy - rad * rad;
  const float disc = b * b - a * c;
  if (disc < 0.0f) { return 1e30f; }
  const float t = (-b - sqrtf(disc)) / a;
  return t > 0.05f ? t : 1e30f;
}

LFX_HD float hit_box(float ox, float oy, float oz, float dx, float dy, float dz,
                     float x0, float x1, float y0, float y1, float z0, float z1)
{
  float tmin = 0.05f, tmax = 1e30f;
  const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz}, lo[3] = {x0, y0, z0}, hi[3] = {x1, y1, z1};
  for (int k = 0; k < 3; k++) {
    if (fabsf(d[k]) < 1e-9f) {
      if (o[k] < lo[k] || o[k] > hi[k]) { return 1e30f; }
    } else {
      float t0 = (lo[k] - o[k]) / d[k], t1 = (hi[k] - o[k]) / d[k];
      if (t0 > t1) { const float t = t0; t0 = t1; t1 = t; }
      tmin = fmaxf(tmin, t0);
      tmax = fminf(tmax, t1);
      if (tmin > tmax) { return 1e30f; }
    }
  }
  return tmin;
}

// range (metres along the ray) of the first surface hit from `pose` in world direction d
LFX_HD float cast(const lfx_synth_spec & s, const Pose & pose, float dx, float dy, float dz)
{
  float Lx, Ly, hf, hc;
  if (s.world == LFX_WORLD_TUNNEL) { Lx = 150.0f; Ly = 3.0f; hf = 2.0f; hc = 4.0f; }
  else { Lx = 10.0f; Ly = 6.0f; hf = 1.8f; hc = 3.0f; }
  float t = slab_exit(pose.px, dx, -Lx, Lx);
  t = fminf(t, slab_exit(pose.py, dy, -Ly, Ly));
  t = fminf(t, slab_exit(0.0f, dz, -hf, hc));
  if (s.world == LFX_WORLD_TUNNEL) {
    // clutter: boxes and poles 2-15 m from the centre line, in front of both walls
    for (int k = 0; k < 10; k++) {
      const float bx = -45.0f + 9.0f * (float)k;
      const float by = (k & 1) ? 1.6f : -1.9f;
      t = fminf(t, hit_box(pose.px, pose.py, 0.0f, dx, dy, dz, bx, bx + 1.2f, by, by + 0.9f, -hf, -hf + 1.5f + 0.2f * (float)(k % 3)));
      t = fminf(t, hit_cylinder(pose.px, pose.py, dx, dy, bx + 4.5f, (k & 1) ? -2.2f : 2.3f, 0.12f));
    }
  } else {
    const float pil[6][3] = {{4.0f, 2.5f, 0.30f}, {-5.0f, -3.0f, 0.25f}, {7.0f, -4.0f, 0.40f},
                             {-7.5f, 3.5f, 0.20f}, {1.5f, -4.5f, 0.15f}, {-2.0f, 4.8f, 0.35f}};
    for (int k = 0; k < 6; k++) {
      t = fminf(t, hit_cylinder(pose.px, pose.py, dx, dy, pil[k][0], pil[k][1], pil[k][2]));
    }
    t = fminf(t, hit_box(pose.px, pose.py, 0.0f, dx, dy, dz, 6.0f, 8.5f, 2.0f, 4.0f, -hf, -0.6f));
    t = fminf(t, hit_box(pose.px, pose.py, 0.0f, dx, dy, dz, -9.0f, -6.5f, -5.5f, -2.5f, -hf, 0.4f));
  }
  return t;
}

// One return. Returns false if the return is dropped (no point emitted).
LFX_HD bool make_point(const lfx_synth_spec & s, uint64_t frame, float az0, uint32_t ring, uint32_t col, Point32 * out)
{
  if (s.dropout_prob > 0.0f) {
    const uint32_t burst = s.dropout_burst >= 1.0f ? (uint32_t)s.dropout_burst : 1u;
    if (u01(draw(s.seed, frame, ring, col / burst, 7)) < s.dropout_prob) { return false; }
  }
  const float two_pi = 6.28318530717958647692f;
  const float step = two_pi / (float)s.n_cols;
  const float jit = (u01(draw(s.seed, frame, ring, col, 1)) - 0.5f) * 0.2f;  // +-10% of a step
  const float az = az0 - ((float)col + jit) * step;                          // clockwise
  const float el = (s.n_rings > 1 ? s.elev_lo_deg + (s.elev_hi_deg - s.elev_lo_deg) * (float)ring / (float)(s.n_rings - 1)
                                  : s.elev_lo_deg) * (two_pi / 360.0f);
  const float ce = cosf(el), se = sinf(el);
  const float sx = ce * cosf(az), sy = ce * sinf(az), sz = se;  // sensor frame
  const Pose pose = pose_of(s, frame);
  const float cyaw = cosf(pose.yaw), syaw = sinf(pose.yaw);
  const float dx = cyaw * sx - syaw * sy, dy = syaw * sx + cyaw * sy, dz = sz;
  float t = cast(s, pose, dx, dy, dz);
  if (s.range_noise > 0.0f) {
    const uint64_t h = draw(s.seed, frame, ring, col, 2);
    const float u1 = fmaxf(u01(h), 1e-7f), u2 = u01(mix64(h));
    t += s.range_noise * sqrtf(-2.0f * logf(u1)) * cosf(two_pi * u2);
  }
  if (s.near_prob > 0.0f && u01(draw(s.seed, frame, ring, col, 3)) < s.near_prob) { t = 0.05f; }
  if (!(t > 0.01f)) { t = 0.01f; }
  out->x = t * sx;
  out->y = t * sy;
  out->z = t * sz;
  out->w = 1.0f;
  out->intensity = 255.0f * u01(draw(s.seed, frame, ring, col, 4));
  out->ring = (uint16_t)ring;
  out->pad0 = 0;
  out->pad1[0] = 0;
  out->pad1[1] = 0;
  return true;
}

LFX_HD float start_azimuth(const lfx_synth_spec & s, uint64_t frame)
{
  return 6.28318530717958647692f * u01(draw(s.seed, frame, 0xFFFFu, 0xFFFFFFu, 5));
}

// ---- the host entry points (lfx_synth_named / lfx_synth_scan_host), shared by liblfx.so and by libsynth.so: the
//      second one is a host-only library, so that a process that only needs scans (bench.py's reference arm, the
//      oracle tests) does not have to map the CUDA product library
inline int named(const char * name, lfx_synth_spec * out)
{
  if (!name || !out) { return LFX_E_BAD_PARAM; }
  lfx_synth_spec s{};
  s.range_noise = 0.005f;
  s.dropout_prob = 0.0f;
  s.dropout_burst = 16.0f;
  s.near_prob = 0.0f;
  s.world = LFX_WORLD_ROOM;
  auto is = [&](const char * n) { const char * a = name; while (*a && *a == *n) { a++; n++; } return *a == 0 && *n == 0; };
  if (is("vlp16")) { s.n_rings = 16; s.n_cols = 1800; s.elev_lo_deg = -15.0f; s.elev_hi_deg = 15.0f; s.seed = 0xC0FFEEull ^ 1; }
  else if (is("hdl32")) { s.n_rings = 32; s.n_cols = 2170; s.elev_lo_deg = -30.67f; s.elev_hi_deg = 10.67f; s.seed = 0xC0FFEEull ^ 2; }
  else if (is("hdl64")) {
    s.n_rings = 64; s.n_cols = 2048; s.elev_lo_deg = 2.0f; s.elev_hi_deg = -24.8f; s.seed = 0xC0FFEEull ^ 3;
    s.world = LFX_WORLD_TUNNEL; s.dropout_prob = 0.01f; s.near_prob = 0.0005f;
  }
  else if (is("os128")) { s.n_rings = 128; s.n_cols = 2048; s.elev_lo_deg = -22.5f; s.elev_hi_deg = 22.5f; s.seed = 0xC0FFEEull ^ 4; }
  else { return LFX_E_BAD_PARAM; }
  *out = s;
  return LFX_OK;
}

inline int scan_host(const lfx_synth_spec * spec, uint64_t frame, void * out, uint32_t * n_points_out)
{
  if (!spec || !out || !n_points_out || spec->n_rings <= 0 || spec->n_cols <= 0 || spec->n_rings > 65535) { return LFX_E_BAD_PARAM; }
  Point32 * p = static_cast<Point32 *>(out);
  const float az0 = start_azimuth(*spec, frame);
  uint32_t n = 0;
  for (int col = 0; col < spec->n_cols; col++) {
    for (int ring = 0; ring < spec->n_rings; ring++) {
      if (make_point(*spec, frame, az0, (uint32_t)ring, (uint32_t)col, &p[n])) { n++; }
    }
  }
  *n_points_out = n;
  return LFX_OK;
}

}  // namespace lfx_synth
#endif  // LFX_SYNTH_H_
