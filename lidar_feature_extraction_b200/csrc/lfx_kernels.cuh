// lfx_kernels.cuh — sm_100a kernels of the extraction path.
//
// Data flow for one batch (all device resident):
//   k_probe_layout + k_extract_sectors (lfx_sector.cuh): the fast path for regular scans; scans that are
//                   not regular, or fail its checks, are flagged and take the general path below
//   k_general_list  compact list of the flagged scans and their ingest tiles
//   k_ring_hist     per 2048-point tile: ring-id histogram           (MakePointIndices, ring.hpp:114-125)
//   k_ring_plan     per scan: stable bucket offsets, ring table (+ RemoveSparseRings, ring.cpp:46-59)
//   k_ring_scatter  per tile: stable scatter of point indices into ring buckets
//   k_probe_rings + k_extract_sectors<indexed> (lfx_sector.cuh): bucketed rings that are rotated monotone
//                   sequences run on the sector kernel through their index list; the rest form the work list of
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

// points per ingest tile: 8192 (fewer, better amortised tile iterations) when the scatter's shared memory allows it,
// i.e. for max_rings <= 2048, else 2048; the kernels are templated on it, the host keeps the choice in the handle
constexpr int TILE_BIG = 8192, TILE_SMALL = 2048;
constexpr int INGEST_THREADS = 256; // 8 warps, 256 consecutive points each
constexpr int MAX_PADDING = 15;     // selection windows live in 16-bit halves
constexpr int MAX_BLOCKS = 64;

enum Counter {
  C_N_WORK = 0, C_WORK_NEXT = 1, C_PACK_NEXT = 2, C_ERR_FLAG = 3, C_ERR_SCAN = 4, C_ERR_RING = 5,
  C_GEN_SCANS = 6,   // scans taking the general path (flagged by k_probe_layout or k_extract_sectors)
  C_GEN_TILES = 7,   // ingest tiles of those scans
  C_N_FAST0 = 8,     // + kidx: entries of the fast-path ring lists (regular scans)
  C_N_FASTX0 = 11,   // + kidx: entries of the indexed ring lists (bucketed rings of flagged scans)
  C_COUNT = 16
};

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
  uint32_t ring16_lo, ring16_hi;   // optional: address of a u16 ring id per point (converter by-product), else 0
                                   // (two words: the descriptor stays 64 bytes, copied with 16-byte cp.async)
};
static_assert(sizeof(ScanDesc) == 64, "ScanDesc is copied as four 16-byte chunks");

__host__ __device__ inline const uint16_t * scan_ring16(const ScanDesc & sd)
{
  return reinterpret_cast<const uint16_t *>((uint64_t)sd.ring16_lo | ((uint64_t)sd.ring16_hi << 32));
}

struct DevParams
{
  int P, B;
  double c_min;   // is_neighbor  <=>  c_min <= cos <= 1.0   (host-derived from acos(c) < theta)
  double c_lo, c_hi;  // c_min * (1 -+ 2^-40): guard band of the division-free link test
  double q_min;   // (double)(float)q > rho  <=>  q >= q_min   (host-derived)
  double q_lo, q_hi;  // q_min * (1 -+ 2^-40)
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
  const uint2 * ring_src;      // [n_scans][max_rings] (first, stride) of rings laid out regularly; stride 0 => idx list
  uint32_t * scan_flags;       // [n_scans] bit 0: the regular-layout hypothesis failed, redo through the general ingest
  const uint2 * work;          // (scan, ring)
  uint32_t * counters;
  uint8_t * labels;
  uint32_t * sorted_src;       // optional
  double * curvature;          // optional
  float4 * stage;              // [total_points]
  int max_rings;
  int cap;                     // ring capacity (multiple of 256); the ring kernel runs cap / 8 threads
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

// ------------------------------------------------------------------ ingest: histogram

// k-th flagged scan owning general tile t: largest k with gen_tile_base[k] <= t
__device__ __forceinline__ int find_gen_scan(const uint32_t * gen_tile_base, int n_gen, uint32_t t)
{
  int lo = 0, hi = n_gen - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (gen_tile_base[mid] <= t) { lo = mid; } else { hi = mid - 1; }
  }
  return lo;
}

// single CTA: list of flagged scans + exclusive prefix of their tile counts
static __global__ void __launch_bounds__(1024)
k_general_list(const ScanDesc * __restrict__ scans, int n_scans, const uint32_t * __restrict__ scan_flags,
               uint32_t * __restrict__ gen_scan, uint32_t * __restrict__ gen_tile_base, uint32_t * __restrict__ tile_owner,
               uint32_t * counters, cudaGraphConditionalHandle cond, int use_cond)
{
  __shared__ uint32_t s_c[1024], s_t[1024];
  const int tid = threadIdx.x;
  uint32_t carry_c = 0, carry_t = 0;
  for (int base = 0; base < n_scans; base += 1024) {
    const int i = base + tid;
    uint32_t f = 0, nt = 0;
    if (i < n_scans && scan_flags[i]) { f = 1; nt = scans[i].n_tiles; }
    s_c[tid] = f; s_t[tid] = nt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t vc = 0, vt = 0;
      if (tid >= off) { vc = s_c[tid - off]; vt = s_t[tid - off]; }
      __syncthreads();
      s_c[tid] += vc; s_t[tid] += vt;
      __syncthreads();
    }
    if (f) {
      const uint32_t k = carry_c + s_c[tid] - 1, tb = carry_t + s_t[tid] - nt;
      gen_scan[k] = (uint32_t)i;
      gen_tile_base[k] = tb;
    }
    carry_c += s_c[1023]; carry_t += s_t[1023];
    __syncthreads();
  }
  if (tid == 0) {
    gen_tile_base[carry_c] = carry_t; counters[C_GEN_SCANS] = carry_c; counters[C_GEN_TILES] = carry_t;
    // inside a CUDA graph the whole general path is the body of an IF node: it only runs when a scan needs it
    if (use_cond) { cudaGraphSetConditional(cond, carry_c != 0 ? 1u : 0u); }
  }
  __syncthreads();
  // general tile -> its entry of gen_scan: one warp per flagged scan, lanes on consecutive tiles (coalesced)
  for (uint32_t k = (uint32_t)tid >> 5; k < carry_c; k += 32) {
    const uint32_t tb = gen_tile_base[k], te = gen_tile_base[k + 1];
    for (uint32_t j = tb + (tid & 31); j < te; j += 32) { tile_owner[j] = k; }
  }
}

// descriptor of general tile t (three dependent loads: owner -> scan -> ScanDesc), fetched one tile ahead by the
// kernels below so that the chain runs while the current tile is processed
struct TileJob { uint32_t scan, tile, first, n_points; uint64_t point_base; const uint8_t * ring_ptr; uint32_t point_step, ring_dt; const uint16_t * src16; };

template<int TILE>
__device__ __forceinline__ TileJob load_tile_job(const ScanDesc * __restrict__ scans, const uint32_t * __restrict__ gen_scan,
                                                 const uint32_t * __restrict__ gen_tile_base, const uint32_t * __restrict__ tile_owner, uint32_t t)
{
  const uint32_t k = tile_owner[t];
  TileJob j;
  j.scan = gen_scan[k];
  const uint32_t in_scan = t - gen_tile_base[k];
  const ScanDesc & sd = scans[j.scan];
  j.tile = sd.tile_base + in_scan;
  j.first = in_scan * TILE;
  j.n_points = sd.n_points;
  j.point_base = sd.point_base;
  j.ring_ptr = sd.data + sd.off_ring;
  j.point_step = sd.point_step;
  j.ring_dt = sd.ring_dt;
  j.src16 = scan_ring16(sd);
  return j;
}

template<int TILE>
__global__ void __launch_bounds__(INGEST_THREADS)
k_ring_hist(const ScanDesc * __restrict__ scans, const uint32_t * __restrict__ gen_scan,
            const uint32_t * __restrict__ gen_tile_base, const uint32_t * __restrict__ tile_owner, uint16_t * __restrict__ ring16,
            uint32_t * __restrict__ tile_hist, int max_rings, uint32_t * counters)
{
  extern __shared__ uint32_t s_hist[];
  constexpr int PER = TILE / INGEST_THREADS;   // 8 points per thread: i = c * 256 + tid
  const int tid = threadIdx.x;
  const uint32_t n_tiles = counters[C_GEN_TILES];
  uint32_t t = blockIdx.x;
  if (t >= n_tiles) { return; }
  TileJob cur = load_tile_job<TILE>(scans, gen_scan, gen_tile_base, tile_owner, t);
  for (; t < n_tiles; t += gridDim.x) {
    const uint32_t tn = t + gridDim.x < n_tiles ? t + gridDim.x : t;
    const TileJob nxt = load_tile_job<TILE>(scans, gen_scan, gen_tile_base, tile_owner, tn);
    uint32_t rg[PER];
#pragma unroll
    for (int c = 0; c < PER; c++) {
      const uint32_t i = cur.first + c * INGEST_THREADS + tid;
      rg[c] = 0xFFFFFFFFu;
      if (i < cur.n_points) {
        rg[c] = cur.src16 ? (uint32_t)cur.src16[i] : load_ring_id(cur.ring_ptr + (size_t)i * cur.point_step, cur.ring_dt);
      }
    }
    __syncthreads();   // the previous tile's histogram has been written out
    for (int r = tid; r < max_rings; r += INGEST_THREADS) { s_hist[r] = 0; }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < PER; c++) {
      if (rg[c] == 0xFFFFFFFFu) { continue; }
      uint32_t ring = rg[c];
      if (ring >= (uint32_t)max_rings) {
        if (atomicExch(&counters[C_ERR_FLAG], LFX_E_CAPACITY) == 0) { counters[C_ERR_SCAN] = cur.scan; counters[C_ERR_RING] = ring; }
        ring = max_rings - 1;
      }
      if (!cur.src16) { ring16[cur.point_base + cur.first + c * INGEST_THREADS + tid] = (uint16_t)ring; }
      atomicAdd(&s_hist[ring], 1u);
    }
    __syncthreads();
    for (int r = tid; r < max_rings; r += INGEST_THREADS) { tile_hist[(size_t)cur.tile * max_rings + r] = s_hist[r]; }
    cur = nxt;
  }
}

// ------------------------------------------------------------------ ingest: plan (one CTA per scan)

static __global__ void __launch_bounds__(256)
k_ring_plan(const ScanDesc * __restrict__ scans, const uint32_t * __restrict__ scan_flags, uint32_t * __restrict__ tile_hist,
            lfx_ring_info * __restrict__ rings, uint2 * __restrict__ ring_src, int max_rings, int padding, int cap)
{
  if (!scan_flags[blockIdx.x]) { return; }  // handled by the fast path
  extern __shared__ uint32_t s_cnt[];  // [max_rings] counts, then [max_rings] offsets
  uint32_t * s_off = s_cnt + max_rings;
  const int scan = blockIdx.x;
  const ScanDesc sd = scans[scan];
  for (int r = threadIdx.x; r < max_rings; r += blockDim.x) {
    uint32_t run = 0;
    // exclusive prefix over the scan's tiles: stable bucket base of every tile. Eight loads are issued before the
    // first store (the compiler cannot prove that the stores do not alias the next loads and would serialise them).
    for (uint32_t t0 = 0; t0 < sd.n_tiles; t0 += 8) {
      uint32_t c[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        c[u] = t0 + u < sd.n_tiles ? tile_hist[(size_t)(sd.tile_base + t0 + u) * max_rings + r] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (t0 + u < sd.n_tiles) { tile_hist[(size_t)(sd.tile_base + t0 + u) * max_rings + r] = run; }
        run += c[u];
      }
    }
    s_cnt[r] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int r = 0; r < max_rings; r++) { s_off[r] = run; run += s_cnt[r]; }
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
    ring_src[(size_t)scan * max_rings + r] = make_uint2(0u, 0u);  // addressed through the bucketed index list
  }
}

// ------------------------------------------------------------------ ingest: stable scatter (one CTA per tile)

// shared memory: [8 warps][R] positions | [R] tile count | [R] tile offset | [R] bucket base | [TILE + R] sorted source
// indices | [TILE + R] their rings (u16) | [8 warps][R] lane probes (bytes). The runs of consecutive rings are one word
// apart more than their lengths: a spinning sensor gives every ring (nearly) the same number of points per tile, a
// multiple of 32 for whole columns, and without the skew 32 lanes with 32 consecutive ring ids would write into ONE bank.
__host__ __device__ inline size_t scatter_smem_bytes(int max_rings, int tile)
{
  return (size_t)(INGEST_THREADS / 32) * max_rings * 5 + (size_t)max_rings * 12 + (size_t)(tile + max_rings) * 6 + 8;
}

template<int TILE>
__global__ void __launch_bounds__(INGEST_THREADS)
k_ring_scatter(const ScanDesc * __restrict__ scans, const uint32_t * __restrict__ gen_scan,
               const uint32_t * __restrict__ gen_tile_base, const uint32_t * __restrict__ tile_owner,
               const uint32_t * __restrict__ counters,
               const uint16_t * __restrict__ ring16, const uint32_t * __restrict__ tile_hist,
               const lfx_ring_info * __restrict__ rings, uint32_t * __restrict__ idx, int max_rings)
{
  extern __shared__ uint32_t s_base[];
  constexpr int WARPS = INGEST_THREADS / 32;
  constexpr int PER_WARP = TILE / WARPS;   // 256 consecutive points
  constexpr int CHUNKS = PER_WARP / 32;    // 8
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, R = max_rings;
  const uint32_t n_tiles = counters[C_GEN_TILES];
  uint32_t t = blockIdx.x;
  if (t >= n_tiles) { return; }
  uint32_t * my = s_base + warp * R;
  uint32_t * tcount = s_base + WARPS * R;
  uint32_t * toff = tcount + R;
  uint32_t * gbase = toff + R;
  uint32_t * s_out = gbase + R;
  uint16_t * s_ring = reinterpret_cast<uint16_t *>(s_out + TILE + R);
  uint8_t * probe = reinterpret_cast<uint8_t *>(s_ring + ((TILE + R + 3) & ~3)) + warp * R;
  __shared__ uint32_t s_total;
  // software pipeline over the CTA's tiles: the descriptor chain (three dependent loads) is fetched two tiles ahead,
  // the ring ids and (for up to 256 rings) the bucket bases one tile ahead, so that none of these latencies is paid
  // between the barriers of a tile
  auto load_ids = [&](const TileJob & j, uint32_t (&ring)[CHUNKS]) {
    const uint16_t * src16 = j.src16 ? j.src16 : ring16 + j.point_base;   // converter by-product, else k_ring_hist's copy
    const uint32_t first = j.first + warp * PER_WARP;
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
      const uint32_t i = first + c * 32 + lane;
      ring[c] = i < j.n_points ? (uint32_t)src16[i] : 0xFFFFFFFFu;
    }
  };
  auto load_base = [&](const TileJob & j) -> uint32_t {
    const int r = threadIdx.x;
    return r < R ? rings[(size_t)j.scan * R + r].offset + tile_hist[(size_t)j.tile * R + r] : 0u;
  };
  auto clamp_t = [&](uint32_t u) { return u < n_tiles ? u : n_tiles - 1; };
  TileJob cur = load_tile_job<TILE>(scans, gen_scan, gen_tile_base, tile_owner, t);
  TileJob nxt = load_tile_job<TILE>(scans, gen_scan, gen_tile_base, tile_owner, clamp_t(t + gridDim.x));
  uint32_t ring[CHUNKS], ring_n[CHUNKS];
  load_ids(cur, ring);
  uint32_t base0 = load_base(cur);
  for (; t < n_tiles; t += gridDim.x) {
    const TileJob nxt2 = load_tile_job<TILE>(scans, gen_scan, gen_tile_base, tile_owner, clamp_t(t + 2 * gridDim.x));
    load_ids(nxt, ring_n);
    const uint32_t base0_n = load_base(nxt);
    const uint32_t first = cur.first + warp * PER_WARP;
    __syncthreads();   // the previous tile's arrays are no longer used
    for (int r = threadIdx.x; r < WARPS * R; r += blockDim.x) { s_base[r] = 0; }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
      if (ring[c] != 0xFFFFFFFFu) { atomicAdd(&my[ring[c]], 1u); }
    }
    __syncthreads();
    // per ring: running prefix over the 8 warps (position inside the ring's run of this tile), the run's length, and
    // where the run goes: the tile's stable base inside the ring bucket
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        const uint32_t c = s_base[w * R + r];
        s_base[w * R + r] = run;
        run += c;
      }
      tcount[r] = run;
      gbase[r] = r == (int)threadIdx.x ? base0 : rings[(size_t)cur.scan * R + r].offset + tile_hist[(size_t)cur.tile * R + r];
    }
    __syncthreads();
    if (warp == 0) {   // exclusive prefix of the run lengths over the rings: the runs' places in the staging array
      uint32_t carry = 0;
      for (int r0 = 0; r0 < R; r0 += 32) {
        const uint32_t v = r0 + lane < R ? tcount[r0 + lane] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) { inc += u; } }
        // (+ one pad word per ring, marked below: the skew that keeps the runs of consecutive rings in different banks)
        if (r0 + lane < R) {
          const uint32_t at = carry + inc - v + (uint32_t)(r0 + lane);
          toff[r0 + lane] = at;
          s_ring[at + v] = 0xFFFFu;
        }
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
      }
      if (lane == 0) { s_total = carry + (uint32_t)R; }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
      const uint32_t i = first + c * 32 + lane;
      const bool valid = ring[c] != 0xFFFFFFFFu;
      // 32 consecutive points of a spinning sensor carry 32 different ring ids: every lane writes its number into
      // its ring's probe slot and finds it again unless another lane has the same ring. Then the stable rank needs
      // no comparison of all lanes with all lanes (match.any, which is slowest exactly when all ids differ).
      if (valid) { probe[ring[c]] = (uint8_t)lane; }
      __syncwarp();
      const bool alone = !valid || probe[ring[c]] == (uint8_t)lane;
      uint32_t lpos = 0;
      if (__all_sync(0xFFFFFFFFu, alone)) {
        if (valid) {
          const uint32_t base = my[ring[c]];
          my[ring[c]] = base + 1u;
          lpos = toff[ring[c]] + base;
        }
      } else {
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ring[c]);
        if (valid) {
          const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
          const uint32_t base = my[ring[c]];
          __syncwarp(peers);
          if (rank == 0) { my[ring[c]] = base + __popc(peers); }
          lpos = toff[ring[c]] + base + rank;
        }
      }
      if (valid) { s_out[lpos] = i; s_ring[lpos] = (uint16_t)ring[c]; }
      __syncwarp();
    }
    __syncthreads();
    // the tile's points, now grouped by ring in stable order, go out run by run: consecutive threads write consecutive
    // words of a ring's bucket (whole 32-byte sectors instead of 2048 scattered 4-byte stores)
    if (R <= 256) {
      // a warp per ring: the run's place is looked up once, the lanes copy consecutive words
      for (int r = warp; r < R; r += WARPS) {
        const uint32_t n = tcount[r];
        const uint32_t * src = s_out + toff[r];
        uint32_t * dst = idx + cur.point_base + gbase[r];
        for (uint32_t k = lane; k < n; k += 32) { dst[k] = src[k]; }
      }
    } else {
      for (uint32_t e = threadIdx.x; e < s_total; e += blockDim.x) {
        const uint32_t r = s_ring[e];
        if (r != 0xFFFFu) { idx[cur.point_base + gbase[r] + (e - toff[r])] = s_out[e]; }
      }
    }
    cur = nxt; nxt = nxt2; base0 = base0_n;
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) { ring[c] = ring_n[c]; }
  }
}

// ------------------------------------------------------------------ packing

// per scan: exclusive prefix of (n_edge, n_surface) over rings ascending -> ring_featoff, totals -> counts
static __global__ void __launch_bounds__(128)
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
static __global__ void __launch_bounds__(1024)
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
static __global__ void __launch_bounds__(256)
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
