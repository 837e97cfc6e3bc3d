// lfx_convert.cuh — the upstream point-type converter on the device.
//
// Replaces PointTypeConverter.callback (point_type_converter/point_type_converter/convert.py:183-212): a raw
// driver PointCloud2 (any field set, any of the eight PointField datatypes, either byte order) becomes the
// deployed 32-byte layout x,y,z,padding,intensity (f32) + ring (u16) (make_fields, convert.py:137-145) with the
// all-zero returns removed (nonzero, convert.py:165-166,192) - in the input's point order.
//
// The struct-format quirks of the reference (effective offsets, positional packing, which fields are tested
// for zero) are resolved on the host into a ConvCloud read plan (lfx_api.cu: conv_make_plan); what runs here
// is byte work: one pass over the input, one CTA per tile of 512 points, the tile staged in shared memory by
// one 1-D bulk copy of the TMA unit (cp.async.bulk + mbarrier), a stable stream compaction (ballot ranks inside
// the tile, decoupled look-back over the tiles of the same cloud, output values decoded while the look-back
// runs) and two 16-byte stores per kept point. Bound: HBM, point_step + 32 * kept bytes
// per point.
#ifndef LFX_CONVERT_CUH_
#define LFX_CONVERT_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#ifndef LFX_CONVERT_TICKET
#define LFX_CONVERT_TICKET 1
#endif

namespace lfxk
{

// tile shape, measured on os128 x 1250 raw clouds (kernel ms): 256 threads x 2 points 7.28, 128 x 4 6.90, 512 x 2 6.96,
// 256 x 4 7.06, 64 x 8 7.01, 128 x 8 7.35, 128 x 2 7.82: four warps per 512-point tile wait least at the tile's barriers
#ifndef LFX_CONV_THREADS
#define LFX_CONV_THREADS 128
#endif
constexpr int CONV_THREADS = LFX_CONV_THREADS;
#ifndef LFX_CONV_PPT
#define LFX_CONV_PPT 4
#endif
constexpr int CONV_PPT = LFX_CONV_PPT;              // points per thread
constexpr int CONV_TILE = CONV_PPT * CONV_THREADS;  // points per CTA
constexpr int CONV_STAGE_MAX_BYTES = 48 * 1024;     // a tile of larger points is decoded straight from global memory
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
  uint8_t fast;           // little-endian, every slot naturally aligned, float32 everywhere but an unsigned ring
  uint8_t pad[1];
};

struct ConvArgs
{
  const ConvCloud * clouds;
  int n_clouds;
  uint32_t n_tiles;
  uint32_t buf_bytes;               // size of one staging buffer (multiple of 128)
  const uint32_t * tile_cloud;      // [n_tiles] owner of each tile
  unsigned long long * tile_state;  // [n_tiles] status << 62 | count
  uint32_t * ticket;
  uint32_t * kept;                  // [n_clouds]
  uint32_t * flags;                 // [n_clouds]
  const uint8_t * out_base;         // first byte of the output buffer (ConvCloud::out points into it)
  uint16_t * ring16;                // by-product: the ring id of output point i of the buffer at [i] (lfx_extract_batch's
                                    // bucketing reads these 2 bytes per point instead of the points' 32-byte sectors)
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

// ---- 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier
__device__ __forceinline__ void conv_mbar_init(uint64_t * bar)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void conv_bulk_load(void * smem_dst, const void * gsrc, uint32_t bytes, uint64_t * bar)
{
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void conv_mbar_wait(uint64_t * bar, uint32_t parity)
{
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(b), "r"(parity) : "memory");
  }
}

// One CTA per tile (CONV_TILE = CONV_PPT * CONV_THREADS consecutive points of one cloud). Thread t owns points t,
// t + CONV_THREADS, ... of the tile: a warp's ballot covers 32 consecutive points and ranks stay stable. (A persistent
// variant with two staging buffers measured 4x slower on B200: with all resident CTAs in lock step every
// look-back has to walk the whole window of concurrently processed tiles.)
// FAST: every cloud of the batch has the common plan (ConvCloud::fast): fields are single 32-bit loads and the
// datatype dispatch disappears; the generic instantiation decodes any plan byte by byte.
template<bool FAST>
__global__ void __launch_bounds__(CONV_THREADS)
k_convert(const ConvArgs a)
{
  extern __shared__ __align__(128) uint8_t conv_raw[];
  __shared__ uint32_t s_tile[1], s_cnt[CONV_PPT * CONV_THREADS / 32];
  __shared__ unsigned long long s_excl;
  __shared__ __align__(8) uint64_t s_bar[1];
  __shared__ ConvCloud s_cc[1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { conv_mbar_init(&s_bar[0]); }
  uint32_t phase[1] = {0u};

  // start staging this CTA's tile into buffer b. The tile comes from an atomic ticket, so a tile's predecessors in the
  // look-back chain have started by construction, whatever order the hardware dispatches CTAs in (blockIdx order is
  // what happens in practice but is not guaranteed: MPS, time slicing, a debugger). -DLFX_CONVERT_TICKET=0 takes
  // tile = blockIdx.x instead (3 % faster on the os128 bench, profiles/r01_summary.md) and relies on in-order dispatch.
  auto fetch = [&](int b) {
#if LFX_CONVERT_TICKET
    if (tid == 0) { s_tile[b] = atomicAdd(a.ticket, 1u); }
    __syncthreads();
    const uint32_t tile = s_tile[b];
    if (tile >= a.n_tiles) { return; }
#else
    const uint32_t tile = blockIdx.x;
    if (tid == 0) { s_tile[b] = tile; }
#endif
    const int c = (int)a.tile_cloud[tile];
    if (tid < (int)(sizeof(ConvCloud) / 4)) { reinterpret_cast<uint32_t *>(&s_cc[b])[tid] = reinterpret_cast<const uint32_t *>(a.clouds + c)[tid]; }
    __syncthreads();
    const ConvCloud & cc = s_cc[b];
    if (!cc.staged) { return; }
    const uint32_t p0 = (tile - cc.tile_base) * CONV_TILE;
    const uint32_t np = min((uint32_t)CONV_TILE, cc.n_points - p0);
    const uint8_t * src = cc.data + (size_t)p0 * cc.point_step;   // 16-byte aligned: the cloud's data pointer is, and CONV_TILE * step
    uint8_t * dst = conv_raw + (size_t)b * a.buf_bytes;
    const uint32_t bytes = np * cc.point_step, vec = bytes & ~15u;
    if (tid == 0 && vec) { conv_bulk_load(dst, src, vec, &s_bar[b]); }
    for (uint32_t i = vec + (uint32_t)tid; i < bytes; i += CONV_THREADS) { dst[i] = src[i]; }
  };

  fetch(0);
  {
    constexpr int cur = 0;
    const uint32_t tile = s_tile[cur];
    if (tile >= a.n_tiles) { return; }
    const ConvCloud & cc = s_cc[cur];
    const int c = (int)a.tile_cloud[tile];
    const uint32_t t_in = tile - cc.tile_base;
    const uint32_t p0 = t_in * CONV_TILE;
    const int np = (int)min((uint32_t)CONV_TILE, cc.n_points - p0);
    const uint32_t step = cc.point_step;
    const uint8_t * base = cc.data + (size_t)p0 * step;
    if (cc.staged) {
      if (((uint32_t)np * step) & ~15u) { conv_mbar_wait(&s_bar[cur], phase[cur]); phase[cur] ^= 1u; }
      base = conv_raw + (size_t)cur * a.buf_bytes;
    }
    __syncthreads();   // the tail bytes written by fetch() are visible
    const bool big = cc.big != 0;
    uint32_t keep = 0, bal[CONV_PPT];
#pragma unroll
    for (int i = 0; i < CONV_PPT; i++) {
      const int p = i * CONV_THREADS + tid;
      bool k = false;
      if (p < np) {
        const uint8_t * mine = base + (size_t)p * step;
        bool zero = true;
#pragma unroll
        for (int f = 0; f < 3; f++) {
          if constexpr (FAST) { zero = zero && (*reinterpret_cast<const uint32_t *>(mine + cc.off[f]) & 0x7FFFFFFFu) == 0; }
          else { zero = zero && conv_is_zero(conv_load(mine + cc.off[f], cc.dt[f], cc.aligned[f] != 0, big), cc.dt[f]); }
        }
        k = !zero;
      }
      bal[i] = __ballot_sync(0xFFFFFFFFu, k);
      keep |= (k ? 1u : 0u) << i;
      if (lane == 0) { s_cnt[i * (CONV_THREADS / 32) + warp] = __popc(bal[i]); }
    }
    __syncthreads();
    // warp 0: stable ranks inside the tile + decoupled look-back over the earlier tiles of the same cloud;
    // the other warps decode their output values meanwhile
    uint32_t rank_base[CONV_PPT];
    if (warp == 0) {
      constexpr int NC = CONV_PPT * CONV_THREADS / 32;
      static_assert(NC <= 32, "one lane per (pass, warp) count");
      const uint32_t mine = lane < NC ? s_cnt[lane] : 0u;
      uint32_t inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) { inc += v; } }
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
      __syncwarp();
      if (lane < NC) { s_cnt[lane] = inc - mine; }
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
    bool overflow = false, range = false;
    uint32_t w[CONV_PPT][6];
    if (cc.packable) {
#pragma unroll
      for (int i = 0; i < CONV_PPT; i++) {
        if (!((keep >> i) & 1u)) { continue; }
        const uint8_t * mine = base + (size_t)(i * CONV_THREADS + tid) * step;
#pragma unroll
        for (int s = 0; s < 5; s++) {
          const int f = 3 + s;
          if constexpr (FAST) {
            const uint32_t b = *reinterpret_cast<const uint32_t *>(mine + cc.off[f]);
            w[i][s] = (b & 0x7FFFFFFFu) > 0x7F800000u ? (b | 0x00400000u) : b;
          } else {
            w[i][s] = conv_to_f32(conv_load(mine + cc.off[f], cc.dt[f], cc.aligned[f] != 0, big), cc.dt[f], overflow);
          }
        }
        if constexpr (FAST) {
          const uint8_t * rp = mine + cc.off[8];
          const uint32_t r = cc.dt[8] == 2 ? (uint32_t)*rp : (cc.dt[8] == 4 ? (uint32_t)*reinterpret_cast<const uint16_t *>(rp) : *reinterpret_cast<const uint32_t *>(rp));
          if (r > 65535u) { range = true; }
          w[i][5] = r & 0xFFFFu;
        } else {
          w[i][5] = conv_to_u16(conv_load(mine + cc.off[8], cc.dt[8], cc.aligned[8] != 0, big), cc.dt[8], range);
        }
      }
    }
    __syncthreads();   // s_excl and the rank bases are published; every thread is done with the staging buffer
    if (cc.packable && keep) {
      const unsigned long long excl = s_excl;
#pragma unroll
      for (int i = 0; i < CONV_PPT; i++) {
        if (!((keep >> i) & 1u)) { continue; }
        rank_base[i] = s_cnt[i * (CONV_THREADS / 32) + warp];
        const uint32_t rank = rank_base[i] + __popc(bal[i] & ((1u << lane) - 1u));
        uint4 * dst = reinterpret_cast<uint4 *>(cc.out + (size_t)(excl + rank) * 32);
        __stcs(dst, make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]));
        __stcs(dst + 1, make_uint4(w[i][4], w[i][5], 0u, 0u));
        a.ring16[(size_t)((cc.out - a.out_base) >> 5) + (size_t)(excl + rank)] = (uint16_t)w[i][5];
      }
      if (overflow || range) { atomicOr(&a.flags[c], (overflow ? CONV_F_OVERFLOW : 0u) | (range ? CONV_F_RING_RANGE : 0u)); }
    }
  }
}

}  // namespace lfxk
#endif  // LFX_CONVERT_CUH_
