// lfx_convert.cuh — the upstream point-type converter on the device.
//
// Replaces PointTypeConverter.callback (point_type_converter/point_type_converter/convert.py:183-212): a raw
// driver PointCloud2 (any field set, any of the eight PointField datatypes, either byte order) becomes the
// deployed 32-byte layout x,y,z,padding,intensity (f32) + ring (u16) (make_fields, convert.py:137-145) with the
// all-zero returns removed (nonzero, convert.py:165-166,192) - in the input's point order.
//
// The struct-format quirks of the reference (effective offsets, positional packing, which fields are tested
// for zero) are resolved on the host into a ConvCloud read plan (lfx_api.cu: conv_make_plan); what runs here
// is byte work: one pass over the input, one CTA per tile of 256 points, the tile staged in shared memory
// with 16-byte loads, a stable stream compaction (ballot ranks inside the tile, decoupled look-back over the
// tiles of the same cloud) and two 16-byte stores per kept point. Bound: HBM, point_step + 32 * kept bytes
// per point.
#ifndef LFX_CONVERT_CUH_
#define LFX_CONVERT_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace lfxk
{

constexpr int CONV_TILE = 256;          // points per CTA = threads per CTA
constexpr int CONV_STAGE_MAX_STEP = 160; // larger points are decoded straight from global memory
constexpr int CONV_N_SLOTS = 9;         // [0..2] zero-test fields, [3..8] the six output slots

// per-cloud flags raised by the kernel (struct.pack's data-dependent failures, convert.py:104-107)
constexpr uint32_t CONV_F_OVERFLOW = 1u;    // finite float64 too large for the 'f' format
constexpr uint32_t CONV_F_RING_RANGE = 2u;  // 'H' format: 0 <= ring <= 65535

struct ConvCloud
{
  const uint8_t * data;   // raw points (device)
  uint8_t * out;          // converted points of this cloud (device, worst-case sized)
  uint32_t n_points, point_step;
  uint32_t tile_base;     // first tile of this cloud
  uint16_t off[CONV_N_SLOTS];   // effective byte offset inside the point (create_point_format, convert.py:69-81)
  uint8_t dt[CONV_N_SLOTS];     // sensor_msgs/PointField datatype id 1..8 (convert.py:40-53)
  uint8_t aligned[CONV_N_SLOTS];// naturally aligned in every point: one typed load instead of byte loads
  uint8_t big;            // is_bigendian
  uint8_t packable;       // exactly six retained fields and an integer ring: points are written
  uint8_t staged;         // tile goes through shared memory
  uint8_t pad[2];
};

struct ConvArgs
{
  const ConvCloud * clouds;
  int n_clouds;
  uint32_t n_tiles;
  unsigned long long * tile_state;  // [n_tiles] status << 62 | count
  uint32_t * ticket;
  uint32_t * kept;                  // [n_clouds]
  uint32_t * flags;                 // [n_clouds]
};

__device__ __forceinline__ int conv_size(uint32_t dt) { return dt <= 2 ? 1 : (dt <= 4 ? 2 : (dt <= 7 ? 4 : 8)); }

// raw bits of one field, in native (little-endian) order
__device__ __forceinline__ unsigned long long conv_load(const uint8_t * p, uint32_t dt, bool aligned, bool big)
{
  const int size = conv_size(dt);
  unsigned long long v = 0;
  if (aligned) {
    if (size == 1) { v = *p; }
    else if (size == 2) { v = *reinterpret_cast<const uint16_t *>(p); }
    else if (size == 4) { v = *reinterpret_cast<const uint32_t *>(p); }
    else { const uint2 w = *reinterpret_cast<const uint2 *>(p); v = (unsigned long long)w.x | ((unsigned long long)w.y << 32); }
    if (big && size > 1) {
      if (size == 2) { v = __byte_perm((uint32_t)v, 0, 0x4401); }
      else if (size == 4) { v = __byte_perm((uint32_t)v, 0, 0x0123); }
      else { v = ((unsigned long long)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123); }
    }
    return v;
  }
  for (int i = 0; i < size; i++) { v |= (unsigned long long)p[i] << (8 * (big ? size - 1 - i : i)); }
  return v;
}

// `value == 0` of the unpacked Python number (nonzero, convert.py:165-166): -0.0 is zero, NaN is not
__device__ __forceinline__ bool conv_is_zero(unsigned long long raw, uint32_t dt)
{
  if (dt == 7) { return (raw & 0x7FFFFFFFull) == 0; }
  if (dt == 8) { return (raw & 0x7FFFFFFFFFFFFFFFull) == 0; }
  return raw == 0;
}

// struct.pack('f', value): the Python number goes to double (exact for every datatype of the table) and is
// narrowed with (float), round to nearest even; a finite double that becomes inf raises OverflowError; NaNs keep
// their sign and top payload bits and come out quiet (x86 cvtsd2ss / cvtss2sd semantics of the reference's host)
__device__ __forceinline__ uint32_t conv_to_f32(unsigned long long raw, uint32_t dt, bool & overflow)
{
  switch (dt) {
    case 1: return __float_as_uint((float)(int8_t)raw);
    case 2: return __float_as_uint((float)(uint8_t)raw);
    case 3: return __float_as_uint((float)(int16_t)raw);
    case 4: return __float_as_uint((float)(uint16_t)raw);
    case 5: return __float_as_uint(__int2float_rn((int32_t)raw));
    case 6: return __float_as_uint(__uint2float_rn((uint32_t)raw));
    case 7: {
      const uint32_t b = (uint32_t)raw;
      return (b & 0x7FFFFFFFu) > 0x7F800000u ? (b | 0x00400000u) : b;
    }
    default: {
      const uint32_t hi = (uint32_t)(raw >> 32);
      if ((raw & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) {
        return (hi & 0x80000000u) | 0x7FC00000u | (uint32_t)((raw >> 29) & 0x003FFFFFull);
      }
      const double d = __longlong_as_double((long long)raw);
      const float f = __double2float_rn(d);
      if (isinf(f) && !isinf(d)) { overflow = true; }
      return __float_as_uint(f);
    }
  }
}

// struct.pack('H', value) of an integer field
__device__ __forceinline__ uint32_t conv_to_u16(unsigned long long raw, uint32_t dt, bool & range)
{
  long long v;
  switch (dt) {
    case 1: v = (int8_t)raw; break;
    case 2: v = (uint8_t)raw; break;
    case 3: v = (int16_t)raw; break;
    case 4: v = (uint16_t)raw; break;
    case 5: v = (int32_t)raw; break;
    default: v = (uint32_t)raw; break;
  }
  if (v < 0 || v > 65535) { range = true; }
  return (uint32_t)v & 0xFFFFu;
}

constexpr unsigned long long CONV_ST_PARTIAL = 1ull << 62, CONV_ST_INCLUSIVE = 2ull << 62, CONV_ST_VALUE = (1ull << 62) - 1;

__global__ void __launch_bounds__(CONV_TILE)
k_convert(const ConvArgs a)
{
  extern __shared__ __align__(16) uint8_t conv_raw[];
  __shared__ uint32_t s_tile, s_cloud, s_warp[CONV_TILE / 32];
  __shared__ unsigned long long s_excl;
  __shared__ ConvCloud s_cc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_tile = atomicAdd(a.ticket, 1u); }   // tiles start in order: look-back never waits on a tile that has not started
  __syncthreads();
  const uint32_t tile = s_tile;
  if (tile >= a.n_tiles) { return; }
  if (tid == 0) {
    int lo = 0, hi = a.n_clouds - 1;   // last cloud whose first tile is <= tile (clouds without points own no tile)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.clouds[mid].tile_base <= tile) { lo = mid; } else { hi = mid - 1; }
    }
    s_cloud = (uint32_t)lo;
  }
  __syncthreads();
  const int c = (int)s_cloud;
  if (tid < (int)(sizeof(ConvCloud) / 4)) { reinterpret_cast<uint32_t *>(&s_cc)[tid] = reinterpret_cast<const uint32_t *>(a.clouds + c)[tid]; }
  __syncthreads();
  const ConvCloud & cc = s_cc;
  const uint32_t t_in = tile - cc.tile_base;
  const uint32_t p0 = t_in * CONV_TILE;
  const int np = (int)min((uint32_t)CONV_TILE, cc.n_points - p0);
  const uint32_t step = cc.point_step;
  const uint8_t * src = cc.data + (size_t)p0 * step;
  const uint8_t * mine = src + (size_t)tid * step;
  if (cc.staged) {
    const uint32_t bytes = (uint32_t)np * step, vec = bytes & ~15u;   // src is 16-byte aligned: data is, and 256 * step
    for (uint32_t i = (uint32_t)tid * 16u; i < vec; i += CONV_TILE * 16u) {
      *reinterpret_cast<uint4 *>(conv_raw + i) = __ldcs(reinterpret_cast<const uint4 *>(src + i));
    }
    for (uint32_t i = vec + (uint32_t)tid; i < bytes; i += CONV_TILE) { conv_raw[i] = src[i]; }
    __syncthreads();
    mine = conv_raw + (size_t)tid * step;
  }
  const bool big = cc.big != 0;
  bool keep = false;
  if (tid < np) {
    bool zero = true;
#pragma unroll
    for (int f = 0; f < 3; f++) { zero = zero && conv_is_zero(conv_load(mine + cc.off[f], cc.dt[f], cc.aligned[f] != 0, big), cc.dt[f]); }
    keep = !zero;
  }
  // stable ranks inside the tile
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
  if (lane == 0) { s_warp[warp] = __popc(bal); }
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < CONV_TILE / 32; w++) { const uint32_t n = s_warp[w]; if (w < warp) { before += n; } total += n; }
  const uint32_t rank = before + __popc(bal & ((1u << lane) - 1u));
  // decoupled look-back over the earlier tiles of the same cloud
  if (warp == 0) {
    unsigned long long * st = a.tile_state + cc.tile_base;
    unsigned long long excl = 0;
    if (t_in == 0) {
      if (lane == 0) { *reinterpret_cast<volatile unsigned long long *>(st) = CONV_ST_INCLUSIVE | total; }
    } else {
      if (lane == 0) { *reinterpret_cast<volatile unsigned long long *>(st + t_in) = CONV_ST_PARTIAL | total; }
      long long p = (long long)t_in - 1;
      for (;;) {
        const long long q = p - lane;
        unsigned long long v = CONV_ST_INCLUSIVE;   // before the cloud's first tile: prefix 0
        if (q >= 0) {
          do { v = *reinterpret_cast<volatile unsigned long long *>(st + q); } while ((v >> 62) == 0);
        }
        const uint32_t incl = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
        const int first = incl ? __ffs(incl) - 1 : 32;   // nearest tile that already knows its inclusive prefix
        unsigned long long add = lane <= first ? (v & CONV_ST_VALUE) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { add += __shfl_xor_sync(0xFFFFFFFFu, add, o); }
        excl += add;
        if (incl) { break; }
        p -= 32;
      }
      if (lane == 0) { *reinterpret_cast<volatile unsigned long long *>(st + t_in) = CONV_ST_INCLUSIVE | (excl + total); }
    }
    if (lane == 0) {
      s_excl = excl;
      if (p0 + (uint32_t)np == cc.n_points) { a.kept[c] = (uint32_t)(excl + total); }
    }
  }
  __syncthreads();
  if (!keep || !cc.packable) { return; }
  bool overflow = false, range = false;
  uint32_t w[6];
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const int f = 3 + s;
    w[s] = conv_to_f32(conv_load(mine + cc.off[f], cc.dt[f], cc.aligned[f] != 0, big), cc.dt[f], overflow);
  }
  w[5] = conv_to_u16(conv_load(mine + cc.off[8], cc.dt[8], cc.aligned[8] != 0, big), cc.dt[8], range);
  uint4 * dst = reinterpret_cast<uint4 *>(cc.out + (size_t)(s_excl + rank) * 32);
  __stcs(dst, make_uint4(w[0], w[1], w[2], w[3]));
  __stcs(dst + 1, make_uint4(w[4], w[5], 0u, 0u));
  if (overflow || range) { atomicOr(&a.flags[c], (overflow ? CONV_F_OVERFLOW : 0u) | (range ? CONV_F_RING_RANGE : 0u)); }
}

}  // namespace lfxk
#endif  // LFX_CONVERT_CUH_
