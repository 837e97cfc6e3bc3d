// lfx_ring.cuh — the per-ring kernel: everything in feature_extraction.cpp:121-151 for one ring, held in
// shared memory by one CTA; a persistent grid walks the (scan, ring) work list.
//
// Shape of the computation (why it looks the way it does; evidence in profiles/):
//  * The path is instruction-issue / latency bound, not DRAM bound. So every thread owns EIGHT consecutive
//    ring positions, keeps the XY-range / curvature sliding windows in registers, and writes each
//    per-point predicate as one BYTE of a bit stream (no ballots, no per-point shared-memory round trips).
//  * Everything that is a neighbourhood predicate (conflict windows, greedy selection, cover fills,
//    occlusion fills, label priority) runs bit-sliced on 32-position words: one thread per word.
//  * Global latency is taken off the critical path: the xyz of ring k+1 are fetched with cp.async
//    (16 B per point, L1-bypassing) into the second half of a double buffer while ring k is processed;
//    the (scan, ring) descriptors are themselves prefetched two rings ahead with cp.async.
//  * All roundings that decide a label are IEEE-exact and uncontracted; divisions are replaced by
//    guard-banded multiplications with an exact fallback inside the band.
#ifndef LFX_RING_CUH_
#define LFX_RING_CUH_

#include "lfx_kernels.cuh"

namespace lfxk
{

constexpr int PTS = 8;  // consecutive ring positions per thread

// bit-stream arrays (one bit per ring position)
enum BitArray {
  A_LINK = 0,  // is_neighbor(i, i+1)                                   neighbor.hpp:44-48
  A_LS,        // link usable by the selection: both ends inside one sector  label.hpp:153-163
  A_TL,        // left occlusion trigger at k = i                        occlusion.hpp:45-53
  A_TRS,       // right occlusion trigger at k = i + 1                   occlusion.hpp:67-75
  A_OOR,       // out of range                                           out_of_range.hpp:36-48
  A_PB,        // parallel beam                                          parallel_beam.hpp:36-51
  A_E,         // edge candidate: inside && curvature >= edge_threshold  label.hpp:81-83
  A_S0,        // surface candidate before the edge pass: inside && curvature <= surface_threshold
  A_SB,        // i is the last position of a sector
  A_XE,        // picked Edge
  A_XS,        // picked Surface
  A_L0, A_L1, A_L2,  // bit planes of the final label
  A_EM, A_SM,  // final label == Edge / == Surface
  A_C0,        // A_C0 + d - 1: curvature(i + d) >= curvature(i), d = 1..P   (key order, ties by index)
  A_COUNT_FIXED = A_C0
};

// what one ring needs, assembled in shared memory by cp.async two rings ahead
struct RingMeta
{
  ScanDesc sd;          // 64 B
  lfx_ring_info info;   // 24 B
  uint2 src;            // (first, stride); stride 0 => bucketed index list
};
static_assert(sizeof(ScanDesc) == 64, "ScanDesc is copied in 16-byte pieces");
static_assert(sizeof(lfx_ring_info) == 24, "lfx_ring_info is copied in 8-byte pieces");
static_assert(sizeof(RingMeta) == 96, "RingMeta layout");

struct RingSmem
{
  float4 * pts[2];          // double-buffered xyz(+pad), bucket order, XOR-swizzled: slot(q) = q ^ ((q >> 3) & 7)
  double * dr;              // XY range by sorted position, padded: slot(p) = p + (p >> 4), halo of 32 in front
  uint32_t * bits;          // [n_arrays][nwords] ; word 0 of each array is a zero pad
  uint32_t * wpre;          // [2][nwords] exclusive prefix of per-word Edge / Surface counts
  uint16_t * perm;          // [cap] sorted position -> bucket position (sort path only)
  RingMeta * meta;          // [3]
  uint2 * items;            // [4]
  int * bnd;                // [MAX_BLOCKS + 1]
  int * misc;               // [16]
  int nwords;               // words per bit array = cap / 32 + 4
};

__host__ __device__ inline int ring_arrays(int P) { return A_COUNT_FIXED + P; }

__host__ __device__ inline size_t ring_smem_bytes(int cap, int P)
{
  const int nwords = cap / 32 + 4;
  size_t b = 0;
  b += (size_t)cap * 16 * 2;                        // pts[2]
  b += (size_t)(cap + cap / 16 + 128) * 8;          // dr (+ halo)
  b += (size_t)ring_arrays(P) * nwords * 4;         // bit streams
  b += (size_t)2 * nwords * 4;                      // wpre
  b += (size_t)cap * 2;                             // perm
  b += 3 * sizeof(RingMeta) + 4 * sizeof(uint2);
  b += (MAX_BLOCKS + 1 + 3) / 4 * 16 + 16 * 4;
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ RingSmem carve_ring(unsigned char * base, int cap, int P)
{
  RingSmem s;
  s.nwords = cap / 32 + 4;
  s.pts[0] = reinterpret_cast<float4 *>(base); base += (size_t)cap * 16;
  s.pts[1] = reinterpret_cast<float4 *>(base); base += (size_t)cap * 16;
  s.dr = reinterpret_cast<double *>(base) + 32; base += (size_t)(cap + cap / 16 + 128) * 8;
  s.meta = reinterpret_cast<RingMeta *>(base); base += 3 * sizeof(RingMeta);
  s.items = reinterpret_cast<uint2 *>(base); base += 4 * sizeof(uint2);
  s.bits = reinterpret_cast<uint32_t *>(base); base += (size_t)ring_arrays(P) * s.nwords * 4;
  s.wpre = reinterpret_cast<uint32_t *>(base); base += (size_t)2 * s.nwords * 4;
  s.bnd = reinterpret_cast<int *>(base); base += (MAX_BLOCKS + 1 + 3) / 4 * 16;
  s.misc = reinterpret_cast<int *>(base); base += 16 * 4;
  s.perm = reinterpret_cast<uint16_t *>(base);
  return s;
}

enum Misc { N_CNT_ASC = 1, N_POS_NONASC = 2, N_POS_ASC = 3, N_SKIP = 4, N_BADRING = 5, N_CNT_DESC = 6, N_POS_NONDESC = 7 };

__device__ __forceinline__ int pslot(int q) { return q ^ ((q >> 3) & 7); }
__device__ __forceinline__ int dslot(int p) { return p + (p >> 4); }

// bit j of result = X(32 w + j + k), 0 <= k < 32
__device__ __forceinline__ uint32_t shr_bits(uint32_t cur, uint32_t next, int k) { return __funnelshift_r(cur, next, k); }
// bit j of result = X(32 w + j - k), 1 <= k < 32
__device__ __forceinline__ uint32_t shl_bits(uint32_t prev, uint32_t cur, int k) { return __funnelshift_r(prev, cur, 32 - k); }

// spread the 8 bits of b into the low bit of 8 bytes (lo: bits 0-3, hi: bits 4-7)
__device__ __forceinline__ void spread8(uint32_t b, uint32_t & lo, uint32_t & hi)
{
  lo = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
  hi = (((b >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
}

__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void * smem_dst, const void * gsrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// OR-reducing barrier over the first `nthreads` threads (multiple of 32) on hardware barrier 1
__device__ __forceinline__ int bar1_or(int pred, int nthreads)
{
  int res;
  asm volatile(
    "{ .reg .pred p, q; setp.ne.s32 q, %1, 0; barrier.red.or.pred p, 1, %2, q; selp.s32 %0, 1, 0, p; }"
    : "=r"(res) : "r"(pred), "r"(nthreads) : "memory");
  return res;
}

// 32-bit (19-bit pseudo angle | 13-bit position) sort key for the non-monotone path
__device__ __forceinline__ uint32_t polar_key19(float x, float y, int q)
{
  return (polar_key(x, y) & 0xFFFFE000u) | (uint32_t)q;
}

__device__ __forceinline__ void bitonic_u32(uint32_t * k, int n2)
{
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo | stride;
        const bool up = (lo & size) == 0;
        const uint32_t a = k[lo], b = k[hi];
        if ((a > b) == up) { k[lo] = b; k[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// exact fallback: bitonic sort of bucket positions with the reference comparator
__device__ __forceinline__ void bitonic_exact(uint16_t * p, int n, int n2, const float4 * pts)
{
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo | stride;
        const bool up = (lo & size) == 0;
        const int a = p[lo], b = p[hi];
        bool a_after_b;
        if (a >= n || b >= n) { a_after_b = a > b; }
        else {
          const float4 pa = pts[pslot(a)], pb = pts[pslot(b)];
          if (polar_less(pb.x, pb.y, pa.x, pa.y)) { a_after_b = true; }
          else if (polar_less(pa.x, pa.y, pb.x, pb.y)) { a_after_b = false; }
          else { a_after_b = a > b; }
        }
        if (a_after_b == up) { p[lo] = (uint16_t)b; p[hi] = (uint16_t)a; }
      }
    }
  }
  __syncthreads();
}

// is_neighbor(i, i+1) for XY points (x0,y0), (x1,y1) with XY ranges r0, r1:
//   acos(dot / (r0 r1)) < theta   <=>   c_min <= RN(dot / (r0 r1)) <= 1      (host-derived c_min)
// decided without the division whenever dot is outside a 2^-39-wide guard band around c_min * r0 r1.
__device__ __forceinline__ bool link_test(double x0, double y0, double x1, double y1, double r0, double r1, const DevParams & prm)
{
  const double dot = __dadd_rn(__dmul_rn(x0, x1), __dmul_rn(y0, y1));
  const double rr = __dmul_rn(r0, r1);
  if (prm.c_min > 0.0 && rr > 0.0 && rr < 1.0e300) {
    if (dot <= rr) {
      if (dot > __dmul_rn(prm.c_hi, rr)) { return true; }
      if (dot < __dmul_rn(prm.c_lo, rr)) { return false; }
    }
  }
  const double c = __ddiv_rn(dot, rr);
  return (c >= prm.c_min) && (c <= 1.0);
}

// (double)(float)(|dr| / r) > rho   <=>   RN(|dr| / r) >= q_min   (host-derived q_min), guard-banded likewise
__device__ __forceinline__ bool ratio_test(double adr, double r, const DevParams & prm)
{
  if (r > 0.0 && r < 1.0e300) {
    if (adr > __dmul_rn(prm.q_hi, r)) { return true; }
    if (adr < __dmul_rn(prm.q_lo, r)) { return false; }
  }
  const float q = __double2float_rn(__ddiv_rn(adr, r));  // parallel_beam.hpp:44-45
  return (double)q > prm.rho;
}

struct OrderMap  // sorted position -> bucket position
{
  int mode;   // 0: start + p, 1: start - p (both mod n)
  int start;
  int n;
  __device__ __forceinline__ int at(int p) const
  {
    int q = mode == 0 ? start + p : start - p;
    if (q >= n) { q -= n; }
    if (q < 0) { q += n; }
    return q;
  }
};

// source point index of bucket position q of a ring
__device__ __forceinline__ uint32_t src_index(const RingArgs & a, const RingMeta & m, uint64_t pos0, int q)
{
  return m.src.y ? m.src.x + (uint32_t)q * m.src.y : a.idx[pos0 + q];
}

// descriptors of work item `w`: the (scan, ring) pair must already sit in s.items[w & 3]
__device__ __forceinline__ void prefetch_meta(const RingArgs & a, const RingSmem & s, uint32_t w)
{
  const uint2 item = s.items[w & 3];
  RingMeta * m = &s.meta[w % 3];
  const char * sd = reinterpret_cast<const char *>(&a.scans[item.x]);
  cp_async16(reinterpret_cast<char *>(&m->sd), sd);
  cp_async16(reinterpret_cast<char *>(&m->sd) + 16, sd + 16);
  cp_async16(reinterpret_cast<char *>(&m->sd) + 32, sd + 32);
  cp_async16(reinterpret_cast<char *>(&m->sd) + 48, sd + 48);
  const size_t k = (size_t)item.x * a.max_rings + item.y;
  const char * ri = reinterpret_cast<const char *>(&a.rings[k]);
  cp_async8(reinterpret_cast<char *>(&m->info), ri);
  cp_async8(reinterpret_cast<char *>(&m->info) + 8, ri + 8);
  cp_async8(reinterpret_cast<char *>(&m->info) + 16, ri + 16);
  cp_async8(&m->src, &a.ring_src[k]);
}

// asynchronous gather of one ring's xyz into a staging buffer (bucket order = source order)
__device__ __forceinline__ void issue_ring_loads(const RingArgs & a, const RingMeta & m, float4 * buf, int tid, int T)
{
  if (m.info.status != LFX_RING_OK) { return; }
  const int n = (int)m.info.count;
  const uint64_t pos0 = m.sd.point_base + m.info.offset;
#pragma unroll
  for (int k = 0; k < PTS; k++) {
    const int q = tid + k * T;
    if (q < n) {
      const uint8_t * p = m.sd.data + (size_t)src_index(a, m, pos0, q) * m.sd.point_step;
      if (m.sd.vec_ok) {
        cp_async16(&buf[pslot(q)], p + m.sd.off_x);
      } else {
        buf[pslot(q)] = make_float4(*reinterpret_cast<const float *>(p + m.sd.off_x), *reinterpret_cast<const float *>(p + m.sd.off_y),
                                    *reinterpret_cast<const float *>(p + m.sd.off_z), 1.0f);
      }
    }
  }
}

template<int PT, int TMAX, int MINB>
__global__ void __launch_bounds__(TMAX, MINB)
k_extract_rings(const RingArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DevParams & prm = a.prm;
  constexpr int PM = PT > 0 ? PT : MAX_PADDING;
  const int P = PT > 0 ? PT : prm.P;
  const int B = prm.B;
  const RingSmem s = carve_ring(smem_raw, a.cap, prm.P);
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  const int NW = s.nwords;
  const int data_words = a.cap / 32;
  const int word_threads = (data_words + 31) & ~31;
  auto arr = [&](int k) -> uint32_t * { return s.bits + k * NW; };
  auto arr_bytes = [&](int k) -> uint8_t * { return reinterpret_cast<uint8_t *>(s.bits + k * NW) + 4; };

  const uint32_t n_work = a.counters[C_N_WORK];
  const uint32_t G = gridDim.x;
  uint32_t w = blockIdx.x;
  if (w >= n_work) { return; }

  // pad words are zero for the whole kernel; data words are rewritten for every ring
  for (int i = tid; i < ring_arrays(prm.P) * NW; i += T) { s.bits[i] = 0; }
  // ---- prologue: descriptors of rings 0, 1 (and the item of ring 2), data of ring 0
  if (tid == 0) {
    s.items[0] = a.work[w];
    if (w + G < n_work) { s.items[1] = a.work[w + G]; }
    if (w + 2 * G < n_work) { s.items[2] = a.work[w + 2 * G]; }
  }
  __syncthreads();
  if (tid == 0) {
    prefetch_meta(a, s, 0);
    if (w + G < n_work) { prefetch_meta(a, s, 1); }
    cp_async_wait_all();
  }
  __syncthreads();
  issue_ring_loads(a, s.meta[0], s.pts[0], tid, T);

  for (uint32_t it = 0; w < n_work; it++, w += G) {
    // ---- top of ring `it`: its xyz (issued one ring ago) and the descriptors of ring it+1 have landed
    if (tid < 16) { s.misc[tid] = (tid == N_POS_NONASC || tid == N_POS_ASC || tid == N_POS_NONDESC) ? -1 : 0; }
    for (int k = tid; k < data_words; k += T) { arr(A_SB)[1 + k] = 0; }
    cp_async_wait_all();
    __syncthreads();
    // descriptors two rings ahead, item three rings ahead, xyz one ring ahead
    if (tid == 0) {
      if (w + 3 * (uint64_t)G < n_work) { cp_async8(&s.items[(it + 3) & 3], &a.work[w + 3 * G]); }
      if (w + 2 * (uint64_t)G < n_work) { prefetch_meta(a, s, it + 2); }
    }
    if (w + G < n_work) { issue_ring_loads(a, s.meta[(it + 1) % 3], s.pts[(it + 1) & 1], tid, T); }

    const RingMeta & meta = s.meta[it % 3];
    float4 * pts = s.pts[it & 1];
    const uint2 item = s.items[it & 3];
    lfx_ring_info * ring_info = &a.rings[(size_t)item.x * a.max_rings + item.y];
    const int n = (int)meta.info.count;
    const uint64_t pos0 = meta.sd.point_base + meta.info.offset;
    const uint32_t status_in = meta.info.status;

    // ---- rings that contribute nothing: sparse (ring.cpp:46-59). Rings over this kernel's capacity are left to
    //      k_extract_rings_big (lfx_big.cuh), which runs after it over the same work list.
    if (status_in != LFX_RING_OK) {
      if (status_in != LFX_RING_TOO_LONG) {
        for (int i = tid; i < n; i += T) {
          a.labels[pos0 + i] = LFX_LABEL_NONE;
          if (a.sorted_src) { a.sorted_src[pos0 + i] = src_index(a, meta, pos0, i); }
          if (a.curvature) { a.curvature[pos0 + i] = 0.0; }
        }
      }
      __syncthreads();
      continue;
    }

    // ---- phase 1: polar-angle order (SortByAtan2, ring.hpp:101-112). One exact comparator evaluation per
    //      cyclic neighbour pair and direction decides whether the ring already is a rotated strictly
    //      ascending (n-1 ascents) or rotated strictly descending (n-1 descents) sequence; ties (equal
    //      angles) go to the sort, which keeps source order among them like the oracle's stable sort. For sources addressed by (first, stride) the ring
    //      field of every point is checked here as well (the lines were just fetched: L2 hits).
    {
      int cnt = 0, cntd = 0, pos_na = -1, pos_nd = -1, bad_ring = 0;
      uint32_t rid[PTS];
      if (meta.src.y) {
#pragma unroll
        for (int m = 0; m < PTS; m++) {
          const int q = tid + m * T;
          rid[m] = item.y;
          if (q < n) {
            rid[m] = load_ring_id(meta.sd.data + (size_t)(meta.src.x + (uint32_t)q * meta.src.y) * meta.sd.point_step + meta.sd.off_ring, meta.sd.ring_dt);
          }
        }
      }
      if (a.force_order_path == 0) {
#pragma unroll
        for (int m = 0; m < PTS; m++) {
          const int q = tid + m * T;
          if (q < n) {
            const int q1 = q + 1 == n ? 0 : q + 1;
            const float4 pa = pts[pslot(q)], pb = pts[pslot(q1)];
            const bool asc = polar_less(pa.x, pa.y, pb.x, pb.y), desc = polar_less(pb.x, pb.y, pa.x, pa.y);
            cnt += asc ? 1 : 0;
            cntd += desc ? 1 : 0;
            if (!asc) { pos_na = q; }    // q grows with m: keeps the largest
            if (!desc) { pos_nd = q; }
          }
        }
      }
      if (meta.src.y) {
#pragma unroll
        for (int m = 0; m < PTS; m++) { bad_ring |= rid[m] != item.y; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        cntd += __shfl_xor_sync(0xFFFFFFFFu, cntd, o);
        pos_na = max(pos_na, __shfl_xor_sync(0xFFFFFFFFu, pos_na, o));
        pos_nd = max(pos_nd, __shfl_xor_sync(0xFFFFFFFFu, pos_nd, o));
        bad_ring |= __shfl_xor_sync(0xFFFFFFFFu, bad_ring, o);
      }
      if (lane == 0) {
        if (cnt) { atomicAdd(&s.misc[N_CNT_ASC], cnt); }
        if (cntd) { atomicAdd(&s.misc[N_CNT_DESC], cntd); }
        if (pos_na >= 0) { atomicMax(&s.misc[N_POS_NONASC], pos_na); }
        if (pos_nd >= 0) { atomicMax(&s.misc[N_POS_NONDESC], pos_nd); }
        if (bad_ring) { s.misc[N_BADRING] = 1; }
      }
      __syncthreads();
    }
    if (s.misc[N_BADRING]) {
      // the (first, stride) hypothesis of k_probe_layout is wrong for this scan: flag it, the general
      // ingest pass redoes the whole scan; nothing of this ring is used
      if (tid == 0) { atomicOr(&a.scan_flags[item.x], 1u); }
      __syncthreads();
      continue;
    }
    OrderMap om;
    om.n = n;
    om.mode = 0; om.start = 0;
    int order_path = a.force_order_path;
    if (order_path == 0) {
      const int cnt_asc = s.misc[N_CNT_ASC], cnt_desc = s.misc[N_CNT_DESC];
      if (n == 1) { om.mode = 0; om.start = 0; }
      else if (cnt_asc == n - 1) { om.mode = 0; om.start = s.misc[N_POS_NONASC] + 1; if (om.start >= n) { om.start -= n; } }
      else if (cnt_desc == n - 1) { om.mode = 1; om.start = s.misc[N_POS_NONDESC]; }
      else { order_path = 1; }
    }
    if (order_path >= 1) {
      int n2 = 32;
      while (n2 < n) { n2 <<= 1; }
      if (order_path == 1) {
        uint32_t * keys = reinterpret_cast<uint32_t *>(s.dr - 32);  // aliases the (not yet written) range array
        for (int i = tid; i < n2; i += T) {
          if (i < n) { const float4 v = pts[pslot(i)]; keys[i] = polar_key19(v.x, v.y, i); } else { keys[i] = 0xFFFFFFFFu; }
        }
        bitonic_u32(keys, n2);
        for (int i = tid; i < n; i += T) { s.perm[i] = (uint16_t)(keys[i] & 0x1FFFu); }
        __syncthreads();
        int bad = 0;  // verify with the exact comparator: no adjacent inversion
        for (int i = tid; i + 1 < n; i += T) {
          const float4 v0 = pts[pslot(s.perm[i])], v1 = pts[pslot(s.perm[i + 1])];
          if (polar_less(v1.x, v1.y, v0.x, v0.y)) { bad = 1; }
        }
        if (__syncthreads_or(bad)) { order_path = 2; }
      }
      if (order_path == 2) {
        uint16_t * pp = reinterpret_cast<uint16_t *>(s.dr - 32);  // n2 <= 8192 entries fit the range array
        for (int i = tid; i < n2; i += T) { pp[i] = (uint16_t)i; }
        bitonic_exact(pp, n, n2, pts);
        for (int i = tid; i < n; i += T) { s.perm[i] = pp[i]; }
        __syncthreads();
      }
      // physically reorder the coordinates so that the map becomes the identity
      float4 g[PTS];
#pragma unroll
      for (int m = 0; m < PTS; m++) {
        const int p = tid + m * T;
        if (p < n) { g[m] = pts[pslot(s.perm[p])]; }
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < PTS; m++) {
        const int p = tid + m * T;
        if (p < n) { pts[pslot(p)] = g[m]; }
      }
      __syncthreads();
    }

    // ---- ring-level preconditions (the reference throws std::invalid_argument, feature_extraction.cpp:154-156)
    bool skip = (n < 2 * P + 1) || (n - 2 * P < B);  // convolution.cpp:39-43, index_range.cpp:35-40
    if (!skip && tid <= B) {  // IndexRange::Boundary index_range.cpp:60-66, evaluated without contraction
      const double sdb = (double)P, edb = (double)(n - P), nb = (double)B, j = (double)tid;
      const double t1 = __dmul_rn(sdb, __dsub_rn(1.0, __ddiv_rn(j, nb)));
      const double t2 = __ddiv_rn(__dmul_rn(edb, j), nb);
      const int b1 = (int)__dadd_rn(t1, t2);
      s.bnd[tid] = b1;
      if (tid >= 1) {  // b1 - 1 is the last position of sector tid - 1
        atomicOr(&arr(A_SB)[1 + ((b1 - 1) >> 5)], 1u << ((b1 - 1) & 31));
      }
    }

    const int p0 = PTS * tid;
    // ---- phase 2: XY range in double (Range, range.hpp:52-56; XYNorm math.hpp:36-39) for the 8 owned points
    float xk[PTS + 1], yk[PTS + 1];
    double rk[PTS];
    if (!skip) {
#pragma unroll
      for (int k = 0; k <= PTS; k++) {
        const int p = p0 + k;
        xk[k] = 0.f; yk[k] = 0.f;
        if (p < n) { const float2 v = *reinterpret_cast<const float2 *>(&pts[pslot(om.at(p))]); xk[k] = v.x; yk[k] = v.y; }
      }
#pragma unroll
      for (int k = 0; k < PTS; k++) {
        const double xd = (double)xk[k], yd = (double)yk[k];
        rk[k] = __dsqrt_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd)));
        s.dr[dslot(p0 + k)] = rk[k];
      }
    }
    __syncthreads();
    if (!skip) {
      int bad = 0;
      if (tid < B && s.bnd[tid + 1] - s.bnd[tid] < 2) { bad = 1; }  // Slice -> NeighborCheckXY ctor, neighbor.hpp:71-75
      if (bad) { s.misc[N_SKIP] = 1; }
    }

    // ---- phase 3: per-point predicates as bytes of bit streams, curvature in a register window
    if (!skip) {
      // XY range of positions p0 - P .. p0 + 8 + 2P - 1  (index j + PM)
      double rw[PTS + 3 * PM];
#pragma unroll
      for (int j = -PM; j < PTS + 2 * PM; j++) {
        if (j >= -P && j < PTS + 2 * P) { rw[j + PM] = (j >= 0 && j < PTS) ? rk[j] : s.dr[dslot(p0 + j)]; }
      }
      uint32_t b_link = 0, b_ls = 0, b_tl = 0, b_trs = 0, b_oor = 0, b_pb = 0, b_e = 0, b_s0 = 0;
      const uint32_t sb = arr_bytes(A_SB)[tid];
      int zero_pair = 0;
#pragma unroll
      for (int k = 0; k < PTS; k++) {
        const int p = p0 + k;
        const double r0 = rw[k + PM], r1 = rw[k + 1 + PM], rm = rw[k - 1 + PM];
        if (p + 1 < n) {
          if (r0 == 0.0 && r1 == 0.0) { zero_pair = 1; }  // CalcRadian throws, math.cpp:40-42
          const bool link = link_test((double)xk[k], (double)yk[k], (double)xk[k + 1], (double)yk[k + 1], r0, r1, prm);
          if (link) {
            b_link |= 1u << k;
            if (p >= P && p + 1 < n - P && !((sb >> k) & 1u)) { b_ls |= 1u << k; }
            if (p < n - P - 1 && r1 > __dadd_rn(r0, prm.d)) { b_tl |= 1u << k; }       // occlusion.hpp:45-53
            if (p + 1 >= P + 1 && r0 > __dadd_rn(r1, prm.d)) { b_trs |= 1u << k; }     // occlusion.hpp:67-75
          }
        }
        if (p < n) {
          if (!(prm.rmin <= r0 && r0 <= prm.rmax)) { b_oor |= 1u << k; }              // out_of_range.hpp:36-48
          if (p >= 1 && p <= n - 2) {                                                    // parallel_beam.hpp:36-51
            if (ratio_test(fabs(__dsub_rn(rm, r0)), r0, prm) && ratio_test(fabs(__dsub_rn(r1, r0)), r0, prm)) { b_pb |= 1u << k; }
          }
        }
      }
      // curvature of positions p0 .. p0 + 8 + P - 1 (CalcCurvature curvature.cpp:44-50: sum left to right
      // from 0.0, centre weight -2P, uncontracted); C_d bits as soon as both operands exist
      double cw[PTS + PM];
      uint32_t b_c[PM];
#pragma unroll
      for (int d = 0; d < PM; d++) { b_c[d] = 0; }
#pragma unroll
      for (int j = 0; j < PTS + PM; j++) {
        if (j < PTS + P) {
          const int p = p0 + j;
          double cv = 0.0;
          if (p >= P && p < n - P) {
            double sum = rw[j - P + PM];
#pragma unroll
            for (int m = -PM + 1; m <= PM; m++) {
              if (m > -P && m <= P) { sum = __dadd_rn(sum, m == 0 ? __dmul_rn(rw[j + PM], prm.center_w) : rw[j + m + PM]); }
            }
            cv = __dmul_rn(sum, sum);
          }
          cw[j] = cv;
          if (j < PTS) {
            if (p >= P && p < n - P) {
              if (cv >= prm.tau_e) { b_e |= 1u << j; }
              if (cv <= prm.tau_s) { b_s0 |= 1u << j; }
            }
            if (a.curvature && p < n) { a.curvature[pos0 + p] = cv; }
          }
#pragma unroll
          for (int d = 1; d <= PM; d++) {
            if (d <= P && j - d >= 0 && j - d < PTS) { if (cv >= cw[j - d]) { b_c[d - 1] |= 1u << (j - d); } }
          }
        }
      }
      arr_bytes(A_LINK)[tid] = (uint8_t)b_link;
      arr_bytes(A_LS)[tid] = (uint8_t)b_ls;
      arr_bytes(A_TL)[tid] = (uint8_t)b_tl;
      arr_bytes(A_TRS)[tid] = (uint8_t)b_trs;
      arr_bytes(A_OOR)[tid] = (uint8_t)b_oor;
      arr_bytes(A_PB)[tid] = (uint8_t)b_pb;
      arr_bytes(A_E)[tid] = (uint8_t)b_e;
      arr_bytes(A_S0)[tid] = (uint8_t)b_s0;
#pragma unroll
      for (int d = 0; d < PM; d++) { if (d < P) { arr_bytes(A_C0 + d)[tid] = (uint8_t)b_c[d]; } }
      if (zero_pair) { s.misc[N_SKIP] = 1; }
    }
    __syncthreads();
    if (!skip && s.misc[N_SKIP]) { skip = true; }

    if (skip) {
      for (int i = tid; i < n; i += T) {
        a.labels[pos0 + i] = LFX_LABEL_NONE;
        if (a.sorted_src) { a.sorted_src[pos0 + i] = src_index(a, meta, pos0, order_path >= 1 ? (int)s.perm[i] : om.at(i)); }
        if (a.curvature) { a.curvature[pos0 + i] = 0.0; }
      }
      if (tid == 0) { ring_info->status = LFX_RING_SKIPPED; ring_info->order_path = order_path; ring_info->n_edge = 0; ring_info->n_surface = 0; }
      __syncthreads();
      continue;
    }

    // ---- phase 4: bit-sliced selection, one thread per 32-position word (first `word_threads` threads,
    //      synchronised on their own hardware barrier).
    //      The greedy walk of label.hpp:85-94 / 124-133 over the (value, index)-sorted order is the
    //      lexicographically-first maximal independent set of the symmetric cover relation
    //      (fill.hpp:101-117 clipped to the sector), i.e. the unique solution of
    //          x_i = cand_i && !exists j in window(i): key(j) before key(i) && x_j
    //      (dependencies are acyclic by key order), reached by chaotic iteration from x = 0.
    if (tid < word_threads) {
      const int wd = tid;
      const bool active = wd < data_words;
      uint32_t gp[PM], gm[PM], sp[PM], sm[PM], vp[PM], vm[PM];
      uint32_t cand_e = 0, cand_s0 = 0;
      uint32_t * XE = arr(A_XE), * XS = arr(A_XS);
      if (active) {
        const uint32_t * LS = arr(A_LS);
        const uint32_t lsm = LS[wd], ls0 = LS[wd + 1], lsp = LS[wd + 2];  // words wd-1, wd, wd+1
        uint32_t v_prev = 0xFFFFFFFFu, v_cur = 0xFFFFFFFFu;              // V_0 = all ones
#pragma unroll
        for (int d = 1; d <= PM; d++) {
          if (d <= P) {
            // V_d(i) = V_{d-1}(i) & LS(i + d - 1): all links between i and i+d usable
            v_cur &= d == 1 ? ls0 : shr_bits(ls0, lsp, d - 1);
            v_prev &= d == 1 ? lsm : shr_bits(lsm, ls0, d - 1);
            const uint32_t * C = arr(A_C0 + d - 1);
            const uint32_t c_prev = C[wd], c_cur = C[wd + 1];
            const uint32_t a_cur = c_cur & v_cur, a_prev = c_prev & v_prev;     // key(i+d) > key(i), in window
            const uint32_t b_cur = ~c_cur & v_cur, b_prev = ~c_prev & v_prev;   // key(i+d) < key(i), in window
            gp[d - 1] = a_cur;                             // edge pass: i+d dominates i
            gm[d - 1] = shl_bits(b_prev, b_cur, d);        // edge pass: i-d dominates i  <=>  key(i) < key(i-d)
            sp[d - 1] = b_cur;                             // surface pass: smaller keys dominate
            sm[d - 1] = shl_bits(a_prev, a_cur, d);
            vp[d - 1] = v_cur;
            vm[d - 1] = shl_bits(v_prev, v_cur, d);
          }
        }
        cand_e = arr(A_E)[wd + 1];
        cand_s0 = arr(A_S0)[wd + 1];
        XE[wd + 1] = 0; XS[wd + 1] = 0;
      }
      bar1_or(0, word_threads);
      // edge pass
      for (;;) {
        int changed = 0;
        if (active) {
          const uint32_t xl = XE[wd], xr = XE[wd + 2];
          const uint32_t x0 = XE[wd + 1];
          uint32_t xc = x0;
          for (int rep = 0; rep < 4; rep++) {   // a few local sweeps against frozen neighbour words
            uint32_t blocked = 0;
#pragma unroll
            for (int d = 1; d <= PM; d++) {
              if (d <= P) { blocked |= (gp[d - 1] & shr_bits(xc, xr, d)) | (gm[d - 1] & shl_bits(xl, xc, d)); }
            }
            const uint32_t xn = cand_e & ~blocked;
            if (xn == xc) { break; }
            xc = xn;
          }
          if (xc != x0) { XE[wd + 1] = xc; changed = 1; }
        }
        if (!bar1_or(changed, word_threads)) { break; }
      }
      uint32_t xe = 0, ce = 0;
      if (active) {
        const uint32_t xl = XE[wd], xc = XE[wd + 1], xr = XE[wd + 2];
        xe = xc; ce = xc;
#pragma unroll
        for (int d = 1; d <= PM; d++) {
          if (d <= P) { ce |= (vp[d - 1] & shr_bits(xc, xr, d)) | (vm[d - 1] & shl_bits(xl, xc, d)); }
        }
      }
      const uint32_t cand_s = cand_s0 & ~ce;   // still Default after the edge pass, label.hpp:125
      // surface pass
      for (;;) {
        int changed = 0;
        if (active) {
          const uint32_t xl = XS[wd], xr = XS[wd + 2];
          const uint32_t x0 = XS[wd + 1];
          uint32_t xc = x0;
          for (int rep = 0; rep < 4; rep++) {
            uint32_t blocked = 0;
#pragma unroll
            for (int d = 1; d <= PM; d++) {
              if (d <= P) { blocked |= (sp[d - 1] & shr_bits(xc, xr, d)) | (sm[d - 1] & shl_bits(xl, xc, d)); }
            }
            const uint32_t xn = cand_s & ~blocked;
            if (xn == xc) { break; }
            xc = xn;
          }
          if (xc != x0) { XS[wd + 1] = xc; changed = 1; }
        }
        if (!bar1_or(changed, word_threads)) { break; }
      }
      // ---- phase 5: covers, occlusion fills, label priority - all per word
      if (active) {
        const uint32_t xl = XS[wd], xc = XS[wd + 1], xr = XS[wd + 2];
        uint32_t cs = xc;
#pragma unroll
        for (int d = 1; d <= PM; d++) {
          if (d <= P) { cs |= (vp[d - 1] & shr_bits(xc, xr, d)) | (vm[d - 1] & shl_bits(xl, xc, d)); }
        }
        // occlusion (occlusion.hpp:37-91): a trigger within P+1 positions whose chain of links reaches i
        const uint32_t * LK = arr(A_LINK), * TL = arr(A_TL), * TRS = arr(A_TRS);
        const uint32_t lkm = LK[wd], lk0 = LK[wd + 1], lkp = LK[wd + 2];
        const uint32_t tlm = TL[wd], tl0 = TL[wd + 1], tr0 = TRS[wd + 1], trp = TRS[wd + 2];
        uint32_t occ = 0, ch = 0xFFFFFFFFu, chr = 0xFFFFFFFFu;
#pragma unroll
        for (int m = 0; m <= PM; m++) {
          if (m <= P) {
            occ |= ch & shl_bits(tlm, tl0, m + 1);                 // TL(i-1-m) & LINK(i-1..i-m)
            ch &= shl_bits(lkm, lk0, m + 1);
            occ |= chr & (m == 0 ? tr0 : shr_bits(tr0, trp, m));   // TRS(i+m) & LINK(i..i+m-1)
            chr &= m == 0 ? lk0 : shr_bits(lk0, lkp, m);
          }
        }
        // final label = ParallelBeam > OutOfRange > Occluded > selection (feature_extraction.cpp:133-138)
        const uint32_t pb = arr(A_PB)[wd + 1], oor = arr(A_OOR)[wd + 1];
        const uint32_t m7 = pb, m5 = oor & ~pb, m6 = occ & ~oor & ~pb, rest = ~(pb | oor | occ);
        const uint32_t m1 = xe & rest, m3 = xc & ~xe & rest, m4 = cs & ~xc & ~xe & rest, m2 = ce & ~xe & ~cs & rest;
        arr(A_L0)[wd + 1] = m7 | m5 | m1 | m3;
        arr(A_L1)[wd + 1] = m7 | m6 | m3 | m2;
        arr(A_L2)[wd + 1] = m7 | m5 | m6 | m4;
        arr(A_EM)[wd + 1] = m1;
        arr(A_SM)[wd + 1] = m3;
        s.wpre[wd] = __popc(m1);
        s.wpre[NW + wd] = __popc(m3);
      }
      bar1_or(0, word_threads);
      // exclusive prefix of the per-word counts: warp 0 for Edge, warp 1 (or warp 0 again) for Surface
      for (int which = tid >> 5; which < 2; which += word_threads >> 5) {
        uint32_t * c = s.wpre + which * NW;
        const int per = (data_words + 31) / 32;
        const int b0 = lane * per;
        uint32_t sum = 0;
        for (int k = 0; k < per; k++) { if (b0 + k < data_words) { sum += c[b0 + k]; } }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) { inc += v; } }
        uint32_t run = inc - sum;
        for (int k = 0; k < per; k++) { if (b0 + k < data_words) { const uint32_t v = c[b0 + k]; c[b0 + k] = run; run += v; } }
        if (lane == 31) { if (which == 0) { ring_info->n_edge = inc; } else { ring_info->n_surface = inc; } }
      }
    }
    __syncthreads();

    // ---- phase 6: outputs. Label bytes (8 per thread), feature points staged (edge ascending from the
    //      ring start, surface descending from the ring end).
    if (p0 < n) {
      uint32_t l0lo, l0hi, l1lo, l1hi, l2lo, l2hi;
      spread8(arr_bytes(A_L0)[tid], l0lo, l0hi);
      spread8(arr_bytes(A_L1)[tid], l1lo, l1hi);
      spread8(arr_bytes(A_L2)[tid], l2lo, l2hi);
      const uint32_t lo = l0lo | (l1lo << 1) | (l2lo << 2), hi = l0hi | (l1hi << 1) | (l2hi << 2);
      uint8_t * dst = a.labels + pos0 + p0;
      if (p0 + PTS <= n && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
      } else {
#pragma unroll
        for (int k = 0; k < PTS; k++) { if (p0 + k < n) { dst[k] = (uint8_t)((k < 4 ? lo >> (8 * k) : hi >> (8 * (k - 4))) & 0xFFu); } }
      }
      const int wq = tid >> 2, sh = (tid & 3) * 8;
      uint32_t em = arr_bytes(A_EM)[tid], smk = arr_bytes(A_SM)[tid];
      if (em | smk) {
        uint32_t re = s.wpre[wq] + __popc(arr(A_EM)[wq + 1] & ((1u << sh) - 1u));
        uint32_t rs = s.wpre[NW + wq] + __popc(arr(A_SM)[wq + 1] & ((1u << sh) - 1u));
        while (em) {
          const int k = __ffs(em) - 1; em &= em - 1;
          float4 v = pts[pslot(om.at(p0 + k))]; v.w = 1.0f;
          a.stage[pos0 + re++] = v;
        }
        while (smk) {
          const int k = __ffs(smk) - 1; smk &= smk - 1;
          float4 v = pts[pslot(om.at(p0 + k))]; v.w = 1.0f;
          a.stage[pos0 + (uint32_t)(n - 1) - rs++] = v;
        }
      }
    }
    if (a.sorted_src) {
      for (int i = tid; i < n; i += T) { a.sorted_src[pos0 + i] = src_index(a, meta, pos0, order_path >= 1 ? (int)s.perm[i] : om.at(i)); }
    }
    if (tid == 0) { ring_info->order_path = order_path; }
    __syncthreads();
  }
  cp_async_wait_all();
}

}  // namespace lfxk
#endif  // LFX_RING_CUH_
