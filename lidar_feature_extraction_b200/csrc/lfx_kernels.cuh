// lfx_kernels.cuh — sm_100a kernels of the extraction path.
//
// Data flow for one batch (all device resident):
//   k_ring_hist     per 2048-point tile: ring-id histogram           (MakePointIndices, ring.hpp:114-125)
//   k_ring_plan     per scan: stable bucket offsets, ring table, work list (+ RemoveSparseRings, ring.cpp:46-59)
//   k_ring_scatter  per tile: stable scatter of point indices into ring buckets
//   k_extract_rings persistent, one CTA per ring at a time: everything in feature_extraction.cpp:121-151
//                   held in shared memory (angle order, range, curvature, link/mask bitfields,
//                   sector-clipped greedy selection, final labels, feature staging)
//   k_feat_offsets_{a,b} + k_pack_copy: canonical (scan, ring asc, index asc) packing of the
//                   edge/surface clouds                              (feature_extraction.cpp:142-151,163-164)
//
// Bit-exactness rules: every double/float operation whose rounding matters is
// written with explicit _rn intrinsics so that nvcc cannot contract it into an FMA.
#ifndef LFX_KERNELS_CUH_
#define LFX_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "lfx.h"

namespace lfxk
{

constexpr int TILE = 2048;          // points per ingest tile
constexpr int INGEST_THREADS = 256; // 8 warps, 256 consecutive points each
constexpr int RING_THREADS = 512;
constexpr int MAX_PADDING = 15;     // selection windows live in 16-bit halves
constexpr int MAX_BLOCKS = 64;

enum Counter { C_N_WORK = 0, C_WORK_NEXT = 1, C_PACK_NEXT = 2, C_ERR_FLAG = 3, C_ERR_SCAN = 4, C_ERR_RING = 5, C_COUNT = 8 };

struct ScanDesc
{
  const uint8_t * data;
  uint64_t point_base;  // first position of this scan in per-point arrays
  uint32_t n_points;
  uint32_t point_step;
  uint32_t off_x, off_y, off_z, off_ring;
  uint32_t ring_dt;
  uint32_t tile_base;   // first ingest tile of this scan
  uint32_t n_tiles;
  uint32_t vec_ok;      // xyz loadable as one aligned float4
};

struct DevParams
{
  int P, B;
  double c_min;   // is_neighbor  <=>  c_min <= cos <= 1.0   (host-derived from acos(c) < theta)
  double d;       // distance_diff_threshold
  double rho;     // parallel_beam_min_range_ratio
  double tau_e, tau_s, rmin, rmax;
  double center_w;  // -2.0 * P
};

struct RingArgs
{
  const ScanDesc * scans;
  const uint32_t * idx;        // [total_points] bucketed source indices
  lfx_ring_info * rings;       // [n_scans][max_rings]
  const uint2 * work;          // (scan, ring)
  uint32_t * counters;
  uint8_t * labels;
  uint32_t * sorted_src;       // optional
  double * curvature;          // optional
  float4 * stage;              // [total_points]
  int max_rings;
  int cap;                     // ring capacity (multiple of 64)
  int cap2;                    // next power of two >= cap
  int force_order_path;
  DevParams prm;
};

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint32_t load_ring_id(const uint8_t * p, uint32_t dt)
{
  if (dt == LFX_RING_U8) { return p[0]; }
  if (dt == LFX_RING_U32) { return *reinterpret_cast<const uint32_t *>(p); }
  return *reinterpret_cast<const uint16_t *>(p);
}

// AHasSmallerPolarAngleThanB<PointXYZIR>, ring.hpp:54-99, float arithmetic without contraction.
__device__ __forceinline__ bool polar_less(float ax, float ay, float bx, float by)
{
  if (ax == bx && ay == by) { return false; }
  const float lena = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
  const float lenb = __fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by));
  if (lena == 0.0f) {
    if (by == 0.0f) { return bx < 0.0f; }
    return by > 0.0f;
  }
  if (lenb == 0.0f) { return ay < 0.0f; }
  if (ay == 0.0f) { return (ax >= 0.0f) && (by >= 0.0f); }
  if (by == 0.0f) { return !((bx >= 0.0f) && (ay >= 0.0f)); }
  if (__fmul_rn(ay, by) > 0.0f) {
    const float det = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
    return det > 0.0f;
  }
  return ay < 0.0f;
}

// Monotone-in-angle 32-bit key (pseudo angle in (-2, 2]); only a sorting accelerator: the order it
// produces is always re-verified with polar_less and replaced by an exact sort if it disagrees.
__device__ __forceinline__ uint32_t polar_key(float x, float y)
{
  const float s = fabsf(x) + fabsf(y);
  const float t = s > 0.0f ? __fdividef(y, s) : 0.0f;  // in [-1, 1]; garbage for inf/NaN is caught by the verify pass
  float k;
  if (x >= 0.0f) { k = t; }
  else if (y >= 0.0f) { k = 2.0f - t; }
  else { k = -2.0f - t; }
  const uint32_t b = __float_as_uint(k);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Bit arrays: word 0 is a zero pad for positions -32..-1, position i lives in word 1 + (i >> 5);
// two zero pad words follow the last data word.
__device__ __forceinline__ uint32_t bits_left32(const uint32_t * w, int i)   // positions [i-32, i); bit 31 <-> i-1
{
  const int q = i >> 5, s = i & 31;
  return __funnelshift_r(w[q], w[q + 1], s);
}
__device__ __forceinline__ uint32_t bits_right32(const uint32_t * w, int i)  // positions [i, i+32); bit 0 <-> i
{
  const int q = (i >> 5) + 1, s = i & 31;
  return __funnelshift_r(w[q], w[q + 1], s);
}
__device__ __forceinline__ uint32_t low_mask(int n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); }

// window of a bit array around i in the "dominator" layout: bit (d-1) <-> i+d, bit 16+(d-1) <-> i-d
__device__ __forceinline__ uint32_t bits_window(const uint32_t * w, int i)
{
  const uint32_t r = bits_right32(w, i + 1) & 0xFFFFu;
  const uint32_t l = __brev(bits_left32(w, i)) & 0xFFFFu;
  return r | (l << 16);
}

// ------------------------------------------------------------------ ingest: histogram

__device__ __forceinline__ int find_scan_of_tile(const ScanDesc * scans, int n_scans, uint32_t tile)
{
  int lo = 0, hi = n_scans - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (scans[mid].tile_base <= tile) { lo = mid; } else { hi = mid - 1; }
  }
  return lo;
}

__global__ void __launch_bounds__(INGEST_THREADS)
k_ring_hist(const ScanDesc * __restrict__ scans, int n_scans, uint16_t * __restrict__ ring16,
            uint32_t * __restrict__ tile_hist, int max_rings, uint32_t * counters)
{
  extern __shared__ uint32_t s_hist[];
  __shared__ int s_scan;
  const uint32_t tile = blockIdx.x;
  if (threadIdx.x == 0) { s_scan = find_scan_of_tile(scans, n_scans, tile); }
  for (int r = threadIdx.x; r < max_rings; r += blockDim.x) { s_hist[r] = 0; }
  __syncthreads();
  const ScanDesc sd = scans[s_scan];
  const uint32_t first = (tile - sd.tile_base) * TILE;
  for (uint32_t k = threadIdx.x; k < TILE; k += blockDim.x) {
    const uint32_t i = first + k;
    if (i < sd.n_points) {
      uint32_t ring = load_ring_id(sd.data + (size_t)i * sd.point_step + sd.off_ring, sd.ring_dt);
      if (ring >= (uint32_t)max_rings) {
        if (atomicExch(&counters[C_ERR_FLAG], LFX_E_CAPACITY) == 0) { counters[C_ERR_SCAN] = s_scan; counters[C_ERR_RING] = ring; }
        ring = max_rings - 1;
      }
      ring16[sd.point_base + i] = (uint16_t)ring;
      atomicAdd(&s_hist[ring], 1u);
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < max_rings; r += blockDim.x) { tile_hist[(size_t)tile * max_rings + r] = s_hist[r]; }
}

// ------------------------------------------------------------------ ingest: plan (one CTA per scan)

__global__ void __launch_bounds__(256)
k_ring_plan(const ScanDesc * __restrict__ scans, uint32_t * __restrict__ tile_hist, lfx_ring_info * __restrict__ rings,
            uint2 * __restrict__ work, uint32_t * counters, int max_rings, int padding, int cap)
{
  extern __shared__ uint32_t s_cnt[];  // [max_rings] counts, then [max_rings] offsets
  uint32_t * s_off = s_cnt + max_rings;
  __shared__ uint32_t s_work_base, s_n_present;
  const int scan = blockIdx.x;
  const ScanDesc sd = scans[scan];
  for (int r = threadIdx.x; r < max_rings; r += blockDim.x) {
    uint32_t run = 0;
    for (uint32_t t = 0; t < sd.n_tiles; t++) {
      uint32_t * h = &tile_hist[(size_t)(sd.tile_base + t) * max_rings + r];
      const uint32_t c = *h;
      *h = run;  // exclusive prefix over the scan's tiles: stable bucket base of this tile
      run += c;
    }
    s_cnt[r] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0, present = 0;
    for (int r = 0; r < max_rings; r++) { s_off[r] = run; run += s_cnt[r]; present += s_cnt[r] ? 1u : 0u; }
    s_n_present = present;
    s_work_base = present ? atomicAdd(&counters[C_N_WORK], present) : 0u;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < max_rings; r += blockDim.x) {
    lfx_ring_info ri;
    ri.count = s_cnt[r];
    ri.offset = s_off[r];
    ri.n_edge = 0;
    ri.n_surface = 0;
    ri.status = ri.count == 0 ? LFX_RING_OK : (ri.count < (uint32_t)(padding + 1) ? LFX_RING_SPARSE : (ri.count > (uint32_t)cap ? LFX_RING_TOO_LONG : LFX_RING_OK));
    ri.order_path = 0;
    rings[(size_t)scan * max_rings + r] = ri;
  }
  if (threadIdx.x == 0 && s_n_present) {
    uint32_t k = s_work_base;
    for (int r = 0; r < max_rings; r++) {
      if (s_cnt[r]) { work[k++] = make_uint2((uint32_t)scan, (uint32_t)r); }
    }
  }
}

// ------------------------------------------------------------------ ingest: stable scatter (one CTA per tile)

__global__ void __launch_bounds__(INGEST_THREADS)
k_ring_scatter(const ScanDesc * __restrict__ scans, int n_scans, const uint16_t * __restrict__ ring16,
               const uint32_t * __restrict__ tile_hist, const lfx_ring_info * __restrict__ rings,
               uint32_t * __restrict__ idx, int max_rings)
{
  extern __shared__ uint32_t s_base[];  // [8 warps][max_rings]
  __shared__ int s_scan;
  constexpr int WARPS = INGEST_THREADS / 32;
  constexpr int PER_WARP = TILE / WARPS;   // 256 consecutive points
  constexpr int CHUNKS = PER_WARP / 32;    // 8
  const uint32_t tile = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { s_scan = find_scan_of_tile(scans, n_scans, tile); }
  for (int r = threadIdx.x; r < WARPS * max_rings; r += blockDim.x) { s_base[r] = 0; }
  __syncthreads();
  const int scan = s_scan;
  const ScanDesc sd = scans[scan];
  const uint32_t first = (tile - sd.tile_base) * TILE + warp * PER_WARP;
  uint32_t * my = s_base + warp * max_rings;

  uint32_t ring[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; c++) {
    const uint32_t i = first + c * 32 + lane;
    ring[c] = i < sd.n_points ? (uint32_t)ring16[sd.point_base + i] : 0xFFFFFFFFu;
    if (ring[c] != 0xFFFFFFFFu) { atomicAdd(&my[ring[c]], 1u); }
  }
  __syncthreads();
  // per ring: running prefix over the 8 warps, seeded with the tile's stable base inside the ring bucket
  for (int r = threadIdx.x; r < max_rings; r += blockDim.x) {
    uint32_t run = rings[(size_t)scan * max_rings + r].offset + tile_hist[(size_t)tile * max_rings + r];
    for (int w = 0; w < WARPS; w++) {
      const uint32_t c = s_base[w * max_rings + r];
      s_base[w * max_rings + r] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < CHUNKS; c++) {
    const uint32_t i = first + c * 32 + lane;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ring[c]);
    if (ring[c] != 0xFFFFFFFFu) {
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      const uint32_t base = my[ring[c]];
      __syncwarp(peers);
      if (rank == 0) { my[ring[c]] = base + __popc(peers); }
      idx[sd.point_base + base + rank] = i;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ the ring kernel

struct RingSmem
{
  double * sr;        // [cap] XY range (f64)            | aliased: u64 sort scratch [cap2] over sr+sc
  double * sc;        // [cap] curvature (f64)
  float * sx, * sy, * sz;   // [cap] loaded (bucket-order) coordinates
  uint32_t * ssrc;    // [cap] source index
  uint32_t * dom;     // [cap] edge dominators inside the conflict window
  uint16_t * perm;    // [cap2] sorted position -> bucket position
  uint8_t * wlr;      // [cap] window half-widths wl | wr << 4
  uint32_t * LINK, * TL, * TRS, * XE, * XS, * CE;  // bit arrays, nwords each
  int * bnd;          // [MAX_BLOCKS + 1]
  int * misc;         // [16]
  int nwords;
};

__host__ __device__ inline size_t ring_smem_bytes(int cap, int cap2)
{
  const int nwords = cap / 32 + 4;
  size_t b = 0;
  b += (size_t)cap * 16;          // sr, sc
  b += (size_t)cap * 12;          // sx, sy, sz
  b += (size_t)cap * 8;           // ssrc, dom
  b += (size_t)cap2 * 2;          // perm
  b += (size_t)cap;               // wlr
  b += (size_t)nwords * 4 * 6;    // bit arrays
  b += (MAX_BLOCKS + 1) * 4 + 16 * 4;
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ RingSmem carve(unsigned char * base, int cap, int cap2)
{
  RingSmem s;
  s.nwords = cap / 32 + 4;
  s.sr = reinterpret_cast<double *>(base); base += (size_t)cap * 8;
  s.sc = reinterpret_cast<double *>(base); base += (size_t)cap * 8;
  s.sx = reinterpret_cast<float *>(base); base += (size_t)cap * 4;
  s.sy = reinterpret_cast<float *>(base); base += (size_t)cap * 4;
  s.sz = reinterpret_cast<float *>(base); base += (size_t)cap * 4;
  s.ssrc = reinterpret_cast<uint32_t *>(base); base += (size_t)cap * 4;
  s.dom = reinterpret_cast<uint32_t *>(base); base += (size_t)cap * 4;
  s.LINK = reinterpret_cast<uint32_t *>(base); base += (size_t)s.nwords * 4;
  s.TL = reinterpret_cast<uint32_t *>(base); base += (size_t)s.nwords * 4;
  s.TRS = reinterpret_cast<uint32_t *>(base); base += (size_t)s.nwords * 4;
  s.XE = reinterpret_cast<uint32_t *>(base); base += (size_t)s.nwords * 4;
  s.XS = reinterpret_cast<uint32_t *>(base); base += (size_t)s.nwords * 4;
  s.CE = reinterpret_cast<uint32_t *>(base); base += (size_t)s.nwords * 4;
  s.bnd = reinterpret_cast<int *>(base); base += (MAX_BLOCKS + 1) * 4;
  s.misc = reinterpret_cast<int *>(base); base += 16 * 4;
  s.perm = reinterpret_cast<uint16_t *>(base); base += (size_t)cap2 * 2;
  s.wlr = reinterpret_cast<uint8_t *>(base);
  return s;
}

enum Misc { M_WORK = 0, M_CNT_A = 1, M_POS_A = 2, M_CNT_D = 3, M_POS_D = 4, M_SKIP = 5, M_FLAG = 6 };

// Block-wide bitonic sort of u64 keys (ascending), n2 a power of two.
__device__ __forceinline__ void bitonic_u64(unsigned long long * k, int n2)
{
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo | stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = k[lo], b = k[hi];
        if ((a > b) == up) { k[lo] = b; k[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// Exact fallback: bitonic sort of bucket positions by the reference comparator (ties: position).
__device__ __forceinline__ void bitonic_exact(uint16_t * p, int n, int n2, const float * sx, const float * sy)
{
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo | stride;
        const bool up = (lo & size) == 0;
        const int a = p[lo], b = p[hi];
        bool a_gt_b;  // "a must come after b"
        if (a >= n || b >= n) { a_gt_b = a > b; }
        else if (polar_less(sx[b], sy[b], sx[a], sy[a])) { a_gt_b = true; }
        else if (polar_less(sx[a], sy[a], sx[b], sy[b])) { a_gt_b = false; }
        else { a_gt_b = a > b; }
        if (a_gt_b == up) { p[lo] = (uint16_t)b; p[hi] = (uint16_t)a; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RING_THREADS, 2)
k_extract_rings(const RingArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RingSmem s = carve(smem_raw, a.cap, a.cap2);
  const int tid = threadIdx.x, lane = tid & 31;
  const DevParams & prm = a.prm;
  const int P = prm.P, B = prm.B;

  for (;;) {
    __syncthreads();
    if (tid == 0) { s.misc[M_WORK] = (int)atomicAdd(&a.counters[C_WORK_NEXT], 1u); }
    __syncthreads();
    const uint32_t w = (uint32_t)s.misc[M_WORK];
    if (w >= a.counters[C_N_WORK]) { break; }
    const uint2 item = a.work[w];
    const ScanDesc sd = a.scans[item.x];
    lfx_ring_info * ring_info = &a.rings[(size_t)item.x * a.max_rings + item.y];
    const int n = (int)ring_info->count;
    const uint64_t pos0 = sd.point_base + ring_info->offset;
    const uint32_t status_in = ring_info->status;

    // ---- rings that contribute nothing: sparse (ring.cpp:46-59) or over capacity
    if (status_in != LFX_RING_OK) {
      for (int i = tid; i < n; i += RING_THREADS) {
        a.labels[pos0 + i] = LFX_LABEL_NONE;
        if (a.sorted_src) { a.sorted_src[pos0 + i] = a.idx[pos0 + i]; }
        if (a.curvature) { a.curvature[pos0 + i] = 0.0; }
      }
      if (status_in == LFX_RING_TOO_LONG && tid == 0) {
        if (atomicExch(&a.counters[C_ERR_FLAG], LFX_E_CAPACITY) == 0) { a.counters[C_ERR_SCAN] = item.x; a.counters[C_ERR_RING] = item.y; }
      }
      continue;
    }

    // ---- phase 0: gather the ring into shared memory (bucket order = source order)
    for (int i = tid; i < s.nwords; i += RING_THREADS) {
      s.LINK[i] = 0; s.TL[i] = 0; s.TRS[i] = 0; s.XE[i] = 0; s.XS[i] = 0; s.CE[i] = 0;
    }
    if (tid < 16 && tid != M_WORK) { s.misc[tid] = 0; }
    for (int i = tid; i < n; i += RING_THREADS) {
      const uint32_t src = a.idx[pos0 + i];
      const uint8_t * p = sd.data + (size_t)src * sd.point_step;
      float x, y, z;
      if (sd.vec_ok) {
        const float4 v = *reinterpret_cast<const float4 *>(p + sd.off_x);
        x = v.x; y = v.y; z = v.z;
      } else {
        x = *reinterpret_cast<const float *>(p + sd.off_x);
        y = *reinterpret_cast<const float *>(p + sd.off_y);
        z = *reinterpret_cast<const float *>(p + sd.off_z);
      }
      s.sx[i] = x; s.sy[i] = y; s.sz[i] = z; s.ssrc[i] = src;
    }
    __syncthreads();

    // ---- phase 1: polar-angle order (SortByAtan2, ring.hpp:101-112)
    // 1a. rotated-monotone test with the exact comparator on cyclic neighbours
    int order_path = a.force_order_path;
    if (order_path == 0) {
      const int nround = (n + 31) & ~31;
      for (int base = 0; base < nround; base += RING_THREADS) {
        const int i = base + tid;
        if (base + (tid & ~31) >= nround) { break; }
        bool descent = false, ascent = false;
        if (i < n) {
          const int j = i + 1 == n ? 0 : i + 1;
          const float ax = s.sx[i], ay = s.sy[i], bx = s.sx[j], by = s.sy[j];
          descent = polar_less(bx, by, ax, ay);
          ascent = polar_less(ax, ay, bx, by);
        }
        const uint32_t bd = __ballot_sync(0xFFFFFFFFu, descent), ba = __ballot_sync(0xFFFFFFFFu, ascent);
        if (lane == 0) {
          if (bd) { atomicAdd(&s.misc[M_CNT_A], __popc(bd)); atomicMax(&s.misc[M_POS_A], base + (tid & ~31) + 31 - __clz(bd)); }
          if (ba) { atomicAdd(&s.misc[M_CNT_D], __popc(ba)); atomicMax(&s.misc[M_POS_D], base + (tid & ~31) + 31 - __clz(ba)); }
        }
      }
      __syncthreads();
      const int cnt_a = s.misc[M_CNT_A], cnt_d = s.misc[M_CNT_D];
      if (cnt_a <= 1) {
        const int start = cnt_a == 1 ? s.misc[M_POS_A] + 1 : 0;
        for (int i = tid; i < n; i += RING_THREADS) { int j = start + i; if (j >= n) { j -= n; } s.perm[i] = (uint16_t)j; }
        order_path = 0;
      } else if (cnt_d == 1) {
        const int b = s.misc[M_POS_D];
        for (int i = tid; i < n; i += RING_THREADS) { int j = b - i; if (j < 0) { j += n; } s.perm[i] = (uint16_t)j; }
        order_path = 0;
      } else {
        order_path = 1;
      }
    }
    if (order_path >= 1) {
      int n2 = 32;
      while (n2 < n) { n2 <<= 1; }
      if (order_path == 1) {
        unsigned long long * keys = reinterpret_cast<unsigned long long *>(s.sr);
        for (int i = tid; i < n2; i += RING_THREADS) {
          keys[i] = i < n ? (((unsigned long long)polar_key(s.sx[i], s.sy[i]) << 32) | (unsigned)i) : ~0ull;
        }
        bitonic_u64(keys, n2);
        for (int i = tid; i < n; i += RING_THREADS) { s.perm[i] = (uint16_t)(keys[i] & 0xFFFFu); }
        __syncthreads();
        // verify with the exact comparator: no adjacent inversion
        int bad = 0;
        for (int i = tid; i + 1 < n; i += RING_THREADS) {
          const int p0 = s.perm[i], p1 = s.perm[i + 1];
          if (polar_less(s.sx[p1], s.sy[p1], s.sx[p0], s.sy[p0])) { bad = 1; }
        }
        if (__syncthreads_or(bad)) { order_path = 2; }
      }
      if (order_path == 2) {
        for (int i = tid; i < n2; i += RING_THREADS) { s.perm[i] = (uint16_t)i; }
        bitonic_exact(s.perm, n, n2, s.sx, s.sy);
      }
    }
    __syncthreads();

    // ---- ring-level preconditions (the reference throws std::invalid_argument, feature_extraction.cpp:154-156)
    bool skip = (n < 2 * P + 1) || (n - 2 * P < B);  // convolution.cpp:39-43, index_range.cpp:35-40
    if (!skip) {
      if (tid <= B) {  // IndexRange::Boundary index_range.cpp:60-66, evaluated without contraction
        const double sdb = (double)P, edb = (double)(n - P), nb = (double)B, j = (double)tid;
        const double t1 = __dmul_rn(sdb, __dsub_rn(1.0, __ddiv_rn(j, nb)));
        const double t2 = __ddiv_rn(__dmul_rn(edb, j), nb);
        s.bnd[tid] = (int)__dadd_rn(t1, t2);
      }
      __syncthreads();
      int bad = 0;
      if (tid < B && s.bnd[tid + 1] - s.bnd[tid] < 2) { bad = 1; }  // Slice -> NeighborCheckXY ctor, neighbor.hpp:71-75
      skip = __syncthreads_or(bad) != 0;
    }

    // ---- phase 2: XY range in double (Range, range.hpp:52-56; XYNorm math.hpp:36-39)
    if (!skip) {
      for (int i = tid; i < n; i += RING_THREADS) {
        const int j = s.perm[i];
        const double xd = (double)s.sx[j], yd = (double)s.sy[j];
        s.sr[i] = __dsqrt_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd)));
      }
      __syncthreads();

      // ---- phase 3: link bits (IsNeighborXY neighbor.hpp:44-48 / CalcRadian math.cpp:34-46) and
      //      occlusion triggers (occlusion.hpp:45-53, 67-75) as bitfields built with ballots
      const int nround = (n + 31) & ~31;
      int zero_pair = 0;
      for (int base = 0; base < nround; base += RING_THREADS) {
        const int i = base + tid;
        if (base + (tid & ~31) >= nround) { break; }
        bool link = false, tl = false, tr = false;
        if (i + 1 < n) {
          const int j0 = s.perm[i], j1 = s.perm[i + 1];
          const double x0 = (double)s.sx[j0], y0 = (double)s.sy[j0], x1 = (double)s.sx[j1], y1 = (double)s.sy[j1];
          const double r0 = s.sr[i], r1 = s.sr[i + 1];
          if (r0 == 0.0 && r1 == 0.0) { zero_pair = 1; }
          const double dot = __dadd_rn(__dmul_rn(x0, x1), __dmul_rn(y0, y1));
          const double c = __ddiv_rn(dot, __dmul_rn(r0, r1));
          link = (c >= prm.c_min) && (c <= 1.0);
          tl = link && (i < n - P - 1) && (r1 > __dadd_rn(r0, prm.d));
          tr = link && (i + 1 >= P + 1) && (r0 > __dadd_rn(r1, prm.d));
        }
        const uint32_t bl = __ballot_sync(0xFFFFFFFFu, link);
        const uint32_t btl = __ballot_sync(0xFFFFFFFFu, tl);
        const uint32_t btr = __ballot_sync(0xFFFFFFFFu, tr);
        if (lane == 0) {
          const int q = 1 + (i >> 5);
          s.LINK[q] = bl; s.TL[q] = btl; s.TRS[q] = btr;  // TRS bit i <-> right trigger at k = i+1
        }
      }
      skip = __syncthreads_or(zero_pair) != 0;  // math.cpp:40-42
    }

    if (skip) {
      for (int i = tid; i < n; i += RING_THREADS) {
        a.labels[pos0 + i] = LFX_LABEL_NONE;
        if (a.sorted_src) { a.sorted_src[pos0 + i] = s.ssrc[s.perm[i]]; }
        if (a.curvature) { a.curvature[pos0 + i] = 0.0; }
      }
      if (tid == 0) { ring_info->status = LFX_RING_SKIPPED; ring_info->order_path = order_path; }
      continue;
    }

    // ---- phase 4: curvature (CalcCurvature curvature.cpp:44-50, Convolution1D convolution.cpp:35-66):
    //      sum left to right from 0.0, centre weight -2P, no contraction
    for (int i = tid; i < n; i += RING_THREADS) {
      double cv = 0.0;
      if (i >= P && i < n - P) {
        double sum = 0.0;
        for (int k = -P; k <= P; k++) {
          const double r = s.sr[i + k];
          sum = __dadd_rn(sum, k == 0 ? __dmul_rn(r, prm.center_w) : r);
        }
        cv = __dmul_rn(sum, sum);
      }
      s.sc[i] = cv;
    }
    __syncthreads();

    // ---- phase 5: per-point conflict window (sector-clipped, chain-aware: fill.hpp:101-117,
    //      label.hpp:153-163) and the set of window neighbours with a larger (curvature, index) key
    for (int i = tid; i < n; i += RING_THREADS) {
      uint32_t dom = 0; int wl = 0, wr = 0;
      if (i >= P && i < n - P) {
        int j = 0;
        while (j + 1 < B && i >= s.bnd[j + 1]) { j++; }
        const int lo = s.bnd[j], hi = s.bnd[j + 1];
        const int lreach = min(P, __clz(~bits_left32(s.LINK, i)));
        const int rreach = min(P, __ffs(~bits_right32(s.LINK, i) | 0x10000u) - 1);
        wl = min(lreach, i - lo);
        wr = min(rreach, hi - 1 - i);
        const double ci = s.sc[i];
        for (int d = 1; d <= wr; d++) { if (s.sc[i + d] >= ci) { dom |= 1u << (d - 1); } }        // key(i+d) > key(i)
        for (int d = 1; d <= wl; d++) { if (s.sc[i - d] > ci) { dom |= 1u << (16 + d - 1); } }    // key(i-d) > key(i)
      }
      s.dom[i] = dom;
      s.wlr[i] = (uint8_t)(wl | (wr << 4));
    }
    __syncthreads();

    // ---- phase 6: greedy selection as the fixpoint x_i = cand_i && no picked dominator in the window.
    //      The greedy walk of label.hpp:85-94 / 124-133 over the (value, index)-sorted order is the
    //      lexicographically-first maximal independent set of the (symmetric) cover relation, which
    //      is the unique solution of this system (dependencies are acyclic by key order).
    const int nround = (n + 31) & ~31;
    // edge pass: descending key => dominators are the larger keys
    for (;;) {
      int changed = 0;
      for (int base = 0; base < nround; base += RING_THREADS) {
        const int i = base + tid;
        if (base + (tid & ~31) >= nround) { break; }
        bool x = false;
        if (i >= P && i < n - P && s.sc[i] >= prm.tau_e) { x = (bits_window(s.XE, i) & s.dom[i]) == 0; }
        const uint32_t bw = __ballot_sync(0xFFFFFFFFu, x);
        if (lane == 0) { const int q = 1 + (i >> 5); if (s.XE[q] != bw) { s.XE[q] = bw; changed = 1; } }
      }
      if (!__syncthreads_or(changed)) { break; }
    }
    // edge cover (Edge or EdgeNeighbor) -> CE
    for (int base = 0; base < nround; base += RING_THREADS) {
      const int i = base + tid;
      if (base + (tid & ~31) >= nround) { break; }
      bool c = false;
      if (i >= P && i < n - P) {
        const int wl = s.wlr[i] & 15, wr = s.wlr[i] >> 4;
        const uint32_t wm = low_mask(wr) | (low_mask(wl) << 16);
        c = ((bits_window(s.XE, i) & wm) != 0) || ((s.XE[1 + (i >> 5)] >> (i & 31)) & 1u);
      }
      const uint32_t bw = __ballot_sync(0xFFFFFFFFu, c);
      if (lane == 0) { s.CE[1 + (i >> 5)] = bw; }
    }
    __syncthreads();
    // surface pass: ascending key => dominators are the smaller keys = window minus the larger ones
    for (;;) {
      int changed = 0;
      for (int base = 0; base < nround; base += RING_THREADS) {
        const int i = base + tid;
        if (base + (tid & ~31) >= nround) { break; }
        bool x = false;
        if (i >= P && i < n - P && !((s.CE[1 + (i >> 5)] >> (i & 31)) & 1u) && s.sc[i] <= prm.tau_s) {
          const int wl = s.wlr[i] & 15, wr = s.wlr[i] >> 4;
          const uint32_t wm = low_mask(wr) | (low_mask(wl) << 16);
          x = (bits_window(s.XS, i) & (wm ^ s.dom[i])) == 0;
        }
        const uint32_t bw = __ballot_sync(0xFFFFFFFFu, x);
        if (lane == 0) { const int q = 1 + (i >> 5); if (s.XS[q] != bw) { s.XS[q] = bw; changed = 1; } }
      }
      if (!__syncthreads_or(changed)) { break; }
    }

    // ---- phase 7: final label (mask overwrites of feature_extraction.cpp:135-138 as a fixed priority),
    //      label bytes out, Edge/Surface points staged (edge ascending from the ring start, surface
    //      descending from the ring end)
    __shared__ uint32_t s_wcnt[2][RING_THREADS / 32];
    uint32_t run_e = 0, run_s = 0;
    for (int base = 0; base < nround; base += RING_THREADS) {
      const int i = base + tid;
      uint8_t label = LFX_LABEL_DEFAULT;
      bool is_e = false, is_s = false;
      if (i < n) {
        if (i >= P && i < n - P) {
          const int wl = s.wlr[i] & 15, wr = s.wlr[i] >> 4;
          const uint32_t wm = low_mask(wr) | (low_mask(wl) << 16);
          const int q = 1 + (i >> 5), sh = i & 31;
          const bool xe = (s.XE[q] >> sh) & 1u, xs = (s.XS[q] >> sh) & 1u, ce = (s.CE[q] >> sh) & 1u;
          const bool cs = (bits_window(s.XS, i) & wm) != 0;
          label = xe ? LFX_LABEL_EDGE : xs ? LFX_LABEL_SURFACE : cs ? LFX_LABEL_SURFACE_NEIGHBOR : ce ? LFX_LABEL_EDGE_NEIGHBOR : LFX_LABEL_DEFAULT;
        }
        // occlusion (occlusion.hpp:37-91): a trigger within P+1 positions whose chain of links reaches i
        const int lreach = min(P, __clz(~bits_left32(s.LINK, i)));
        const int rreach = min(P, __ffs(~bits_right32(s.LINK, i) | 0x10000u) - 1);
        const bool occ_l = (bits_left32(s.TL, i) >> (31 - lreach)) != 0;
        const bool occ_r = (bits_right32(s.TRS, i) & low_mask(rreach + 1)) != 0;
        if (occ_l || occ_r) { label = LFX_LABEL_OCCLUDED; }
        const double r = s.sr[i];
        if (!(prm.rmin <= r && r <= prm.rmax)) { label = LFX_LABEL_OUT_OF_RANGE; }  // out_of_range.hpp:36-48
        if (i >= 1 && i <= n - 2) {  // parallel_beam.hpp:36-51: ratios narrowed to float, compared in double
          const float q1 = __double2float_rn(__ddiv_rn(fabs(__dsub_rn(s.sr[i - 1], r)), r));
          const float q2 = __double2float_rn(__ddiv_rn(fabs(__dsub_rn(s.sr[i + 1], r)), r));
          if ((double)q1 > prm.rho && (double)q2 > prm.rho) { label = LFX_LABEL_PARALLEL_BEAM; }
        }
        a.labels[pos0 + i] = label;
        if (a.sorted_src) { a.sorted_src[pos0 + i] = s.ssrc[s.perm[i]]; }
        if (a.curvature) { a.curvature[pos0 + i] = s.sc[i]; }
        is_e = label == LFX_LABEL_EDGE;
        is_s = label == LFX_LABEL_SURFACE;
      }
      const uint32_t be = __ballot_sync(0xFFFFFFFFu, is_e), bs = __ballot_sync(0xFFFFFFFFu, is_s);
      if (lane == 0) { s_wcnt[0][tid >> 5] = __popc(be); s_wcnt[1][tid >> 5] = __popc(bs); }
      __syncthreads();
      uint32_t pre_e = 0, pre_s = 0, tot_e = 0, tot_s = 0;
      for (int w2 = 0; w2 < RING_THREADS / 32; w2++) {
        const uint32_t ce = s_wcnt[0][w2], cs = s_wcnt[1][w2];
        if (w2 < (tid >> 5)) { pre_e += ce; pre_s += cs; }
        tot_e += ce; tot_s += cs;
      }
      if (is_e || is_s) {
        const int j = s.perm[i];
        const float4 v = make_float4(s.sx[j], s.sy[j], s.sz[j], 1.0f);
        const uint32_t lt = (1u << lane) - 1u;
        if (is_e) { a.stage[pos0 + run_e + pre_e + __popc(be & lt)] = v; }
        else { a.stage[pos0 + (uint32_t)(n - 1) - (run_s + pre_s + __popc(bs & lt))] = v; }
      }
      run_e += tot_e; run_s += tot_s;
      __syncthreads();
    }
    if (tid == 0) { ring_info->n_edge = run_e; ring_info->n_surface = run_s; ring_info->order_path = order_path; }
  }
}

// ------------------------------------------------------------------ packing

// per scan: exclusive prefix of (n_edge, n_surface) over rings ascending -> ring_featoff, totals -> counts
__global__ void __launch_bounds__(128)
k_feat_offsets_a(const lfx_ring_info * __restrict__ rings, uint2 * __restrict__ ring_featoff,
                 uint32_t * __restrict__ counts, int max_rings)
{
  __shared__ uint32_t s_e[128], s_s[128];
  const int scan = blockIdx.x, tid = threadIdx.x;
  uint32_t carry_e = 0, carry_s = 0;
  for (int base = 0; base < max_rings; base += 128) {
    const int r = base + tid;
    uint32_t e = 0, sf = 0;
    if (r < max_rings) { const lfx_ring_info ri = rings[(size_t)scan * max_rings + r]; e = ri.n_edge; sf = ri.n_surface; }
    s_e[tid] = e; s_s[tid] = sf;
    __syncthreads();
    for (int off = 1; off < 128; off <<= 1) {
      uint32_t ve = 0, vs = 0;
      if (tid >= off) { ve = s_e[tid - off]; vs = s_s[tid - off]; }
      __syncthreads();
      s_e[tid] += ve; s_s[tid] += vs;
      __syncthreads();
    }
    if (r < max_rings) { ring_featoff[(size_t)scan * max_rings + r] = make_uint2(carry_e + s_e[tid] - e, carry_s + s_s[tid] - sf); }
    carry_e += s_e[127]; carry_s += s_s[127];
    __syncthreads();
  }
  if (tid == 0) { counts[2 * scan] = carry_e; counts[2 * scan + 1] = carry_s; }
}

// single CTA: exclusive prefix over scans -> offsets[(n_scans+1)][2]
__global__ void __launch_bounds__(1024)
k_feat_offsets_b(const uint32_t * __restrict__ counts, uint32_t * __restrict__ offsets, int n_scans)
{
  __shared__ uint32_t s_e[1024], s_s[1024];
  const int tid = threadIdx.x;
  uint32_t carry_e = 0, carry_s = 0;
  for (int base = 0; base < n_scans; base += 1024) {
    const int i = base + tid;
    uint32_t e = 0, sf = 0;
    if (i < n_scans) { e = counts[2 * i]; sf = counts[2 * i + 1]; }
    s_e[tid] = e; s_s[tid] = sf;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t ve = 0, vs = 0;
      if (tid >= off) { ve = s_e[tid - off]; vs = s_s[tid - off]; }
      __syncthreads();
      s_e[tid] += ve; s_s[tid] += vs;
      __syncthreads();
    }
    if (i < n_scans) { offsets[2 * i] = carry_e + s_e[tid] - e; offsets[2 * i + 1] = carry_s + s_s[tid] - sf; }
    carry_e += s_e[1023]; carry_s += s_s[1023];
    __syncthreads();
  }
  if (tid == 0) { offsets[2 * n_scans] = carry_e; offsets[2 * n_scans + 1] = carry_s; }
}

// persistent: one work item (scan, ring) per CTA iteration; staged -> final concatenated clouds
__global__ void __launch_bounds__(256)
k_pack_copy(const uint2 * __restrict__ work, uint32_t * counters, const ScanDesc * __restrict__ scans,
            const lfx_ring_info * __restrict__ rings, const uint2 * __restrict__ ring_featoff,
            const uint32_t * __restrict__ offsets, const float4 * __restrict__ stage,
            float4 * __restrict__ edge, float4 * __restrict__ surface, int max_rings)
{
  __shared__ uint32_t s_w;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) { s_w = atomicAdd(&counters[C_PACK_NEXT], 1u); }
    __syncthreads();
    const uint32_t w = s_w;
    if (w >= counters[C_N_WORK]) { break; }
    const uint2 item = work[w];
    const lfx_ring_info ri = rings[(size_t)item.x * max_rings + item.y];
    if (ri.n_edge + ri.n_surface == 0) { continue; }
    const uint2 fo = ring_featoff[(size_t)item.x * max_rings + item.y];
    const uint64_t pos0 = scans[item.x].point_base + ri.offset;
    float4 * de = edge + offsets[2 * item.x] + fo.x;
    float4 * ds = surface + offsets[2 * item.x + 1] + fo.y;
    for (uint32_t k = threadIdx.x; k < ri.n_edge; k += blockDim.x) { de[k] = stage[pos0 + k]; }
    for (uint32_t k = threadIdx.x; k < ri.n_surface; k += blockDim.x) { ds[k] = stage[pos0 + ri.count - 1 - k]; }
  }
}

}  // namespace lfxk
#endif  // LFX_KERNELS_CUH_
