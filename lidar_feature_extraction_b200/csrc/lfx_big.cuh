// lfx_big.cuh — k_extract_rings_big: the per-ring pipeline without any size limit, for what the two on-chip kernels
// do not hold: rings longer than lfx_options.max_ring_points (an HDL-64 at 5 Hz, a 4096-column sweep), paddings above
// MAX_PADDING, more than MAX_BLOCKS sectors. The reference accepts all of these (hyper_parameter.hpp:45-53 only asserts
// "> 0", index_range.cpp:32-66 has no upper bound), so the library must not refuse them; it runs them here, one CTA per
// ring, straight from global memory. This is the correctness net, not a fast path: every step is the reference's step in
// its plainest data-parallel form (the same bit-sliced fix-point as the other two kernels, one byte per position).
//
// Scratch is the ring's own slice [pos0, pos0 + n) of per-point arrays the batch already owns:
//   idx    the ring's bucket of source indices (MakePointIndices, ring.hpp:114-125): sorted IN PLACE by polar angle
//   stage  16 bytes per position: (XY range, curvature) as two doubles until the features are written over them
//   labels one byte of flags per position (below) until the final labels replace them
//   tmp16  final label of each position before it is copied over the flags (the dead u16 ring-id copy of the ingest)
#ifndef LFX_BIG_CUH_
#define LFX_BIG_CUH_

#include "lfx_ring.cuh"
#include "lfx_sector.cuh"

namespace lfxk
{

constexpr int BIG_THREADS = 256;

enum BigFlag : uint32_t {
  F_LS = 1u,       // link(p, p+1) usable by the selection: both ends inside one sector (fill.hpp:101-117 on the sector's slice)
  F_XE = 2u,       // picked as Edge
  F_CE = 4u,       // Edge or inside an Edge pick's cover
  F_XS = 8u,       // picked as Surface
  F_CS = 16u,      // Surface or inside a Surface pick's cover
  F_CAND_E = 32u,  // curvature >= edge_threshold, inside a sector (label.hpp:81-83)
  F_CAND_S = 64u,  // curvature <= surface_threshold, inside a sector (label.hpp:120-122)
  F_LINK = 128u    // IsNeighborXY(p, p+1) (neighbor.hpp:44-48)
};

struct BigArgs
{
  const ScanDesc * scans;
  uint32_t * idx;
  lfx_ring_info * rings;
  const uint2 * work;
  uint32_t * counters;
  uint8_t * labels;
  uint32_t * sorted_src;   // optional
  double * curvature;      // optional
  float4 * stage;
  uint16_t * tmp16;
  int max_rings;
  int all;                 // 1: every work item (the on-chip ring kernel is not launched), 0: only LFX_RING_TOO_LONG
  DevParams prm;
};

__device__ __forceinline__ float2 big_xy(const ScanDesc & sd, uint32_t src)
{
  const uint8_t * q = sd.data + (size_t)src * sd.point_step;
  return make_float2(*reinterpret_cast<const float *>(q + sd.off_x), *reinterpret_cast<const float *>(q + sd.off_y));
}

// a before b in the ring's order: smaller polar angle (ring.hpp:54-99), equal angles in source order (stable sort)
__device__ __forceinline__ bool big_before(const ScanDesc & sd, uint32_t a, uint32_t b)
{
  const float2 pa = big_xy(sd, a), pb = big_xy(sd, b);
  if (polar_less(pa.x, pa.y, pb.x, pb.y)) { return true; }
  if (polar_less(pb.x, pb.y, pa.x, pa.y)) { return false; }
  return a < b;
}

__device__ __forceinline__ void big_reverse(uint32_t * v, int lo, int hi)   // [lo, hi); callers synchronise
{
  const int half = (hi - lo) >> 1;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const uint32_t a = v[lo + i], b = v[hi - 1 - i];
    v[lo + i] = b; v[hi - 1 - i] = a;
  }
}

__global__ void __launch_bounds__(BIG_THREADS)
k_extract_rings_big(const BigArgs a)
{
  __shared__ int s_cnt[4];
  __shared__ uint32_t s_scan[2][BIG_THREADS];
  const DevParams & prm = a.prm;
  const int P = prm.P, B = prm.B;
  const int tid = threadIdx.x, T = blockDim.x;
  const uint32_t n_work = a.counters[C_N_WORK];
  for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
    const uint2 item = a.work[w];
    lfx_ring_info * ring_info = &a.rings[(size_t)item.x * a.max_rings + item.y];
    const lfx_ring_info info = *ring_info;
    if (!a.all && info.status != LFX_RING_TOO_LONG) { continue; }
    if (info.count == 0) { continue; }
    const ScanDesc sd = a.scans[item.x];
    const int n = (int)info.count;
    const uint64_t pos0 = sd.point_base + info.offset;
    uint32_t * ord = a.idx + pos0;
    volatile uint8_t * fl = a.labels + pos0;
    double2 * rc = reinterpret_cast<double2 *>(a.stage + pos0);
    __syncthreads();   // (s_cnt and s_scan of the previous ring)

    // ---- rings that contribute nothing: RemoveSparseRings (ring.cpp:46-59)
    if (info.status == LFX_RING_SPARSE) {
      for (int i = tid; i < n; i += T) {
        fl[i] = LFX_LABEL_NONE;
        if (a.sorted_src) { a.sorted_src[pos0 + i] = ord[i]; }
        if (a.curvature) { a.curvature[pos0 + i] = 0.0; }
      }
      continue;
    }

    // ---- polar-angle order (SortByAtan2, ring.hpp:101-112): a rotated strictly monotone ring (spinning sensor) is
    //      rotated / reversed in place; anything else is sorted with the exact comparator (normalised bitonic network:
    //      every exchange ascending, so the virtual +inf padding beyond n never moves and is simply skipped)
    if (tid < 4) { s_cnt[tid] = tid < 2 ? 0 : -1; }
    __syncthreads();
    {
      int cnt = 0, cntd = 0, pna = -1, pnd = -1;
      for (int q = tid; q < n; q += T) {
        const float2 pa = big_xy(sd, ord[q]), pb = big_xy(sd, ord[q + 1 == n ? 0 : q + 1]);
        const bool asc = polar_less(pa.x, pa.y, pb.x, pb.y), desc = polar_less(pb.x, pb.y, pa.x, pa.y);
        cnt += asc ? 1 : 0; cntd += desc ? 1 : 0;
        if (!asc) { pna = q; }
        if (!desc) { pnd = q; }
      }
      if (cnt) { atomicAdd(&s_cnt[0], cnt); }
      if (cntd) { atomicAdd(&s_cnt[1], cntd); }
      if (pna >= 0) { atomicMax(&s_cnt[2], pna); }
      if (pnd >= 0) { atomicMax(&s_cnt[3], pnd); }
    }
    __syncthreads();
    int order_path = 0;
    {
      const int cnt_asc = s_cnt[0], cnt_desc = s_cnt[1], pna = s_cnt[2], pnd = s_cnt[3];
      if (n == 1) {
      } else if (cnt_asc == n - 1) {              // sorted position p = slot (start + p) mod n: rotate left by start
        int start = pna + 1;
        if (start >= n) { start -= n; }
        if (start) {
          big_reverse(ord, 0, start); big_reverse(ord, start, n);
          __syncthreads();
          big_reverse(ord, 0, n);
        }
      } else if (cnt_desc == n - 1) {             // sorted position p = slot (start - p) mod n
        const int start = pnd;
        big_reverse(ord, 0, n);                   // slot q -> n - 1 - q: the smallest angle sits at n - 1 - start
        __syncthreads();
        const int rot = n - 1 - start;
        if (rot) {
          big_reverse(ord, 0, rot); big_reverse(ord, rot, n);
          __syncthreads();
          big_reverse(ord, 0, n);
        }
      } else {
        order_path = 2;
        unsigned int n2 = 2;
        while (n2 < (unsigned int)n) { n2 <<= 1; }
        for (unsigned int k = 2; k <= n2; k <<= 1) {
          for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (unsigned int i = tid; i < (unsigned int)n; i += T) {
              const unsigned int l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
              if (l > i && l < (unsigned int)n) {
                const uint32_t va = ord[i], vb = ord[l];
                if (big_before(sd, vb, va)) { ord[i] = vb; ord[l] = va; }
              }
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- ring-level preconditions (the reference throws std::invalid_argument, feature_extraction.cpp:154-156)
    bool skip = (n < 2 * P + 1) || (n - 2 * P < B);   // convolution.cpp:39-43, index_range.cpp:35-40
    if (tid == 0) { s_cnt[0] = 0; }
    __syncthreads();
    if (!skip) {
      for (int j = tid; j < B; j += T) {                // Slice -> NeighborCheckXY ctor, neighbor.hpp:71-75
        if (sector_bound(P, n, B, j + 1) - sector_bound(P, n, B, j) < 2) { s_cnt[0] = 1; }
      }
    }
    // ---- XY range (Range, range.hpp:52-56; XYNorm math.hpp:36-39); flags start empty
    for (int p = tid; p < n; p += T) {
      const float2 v = big_xy(sd, ord[p]);
      const double xd = (double)v.x, yd = (double)v.y;
      rc[p].x = __dsqrt_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd)));
      fl[p] = 0;
    }
    __syncthreads();
    if (!skip) {
      // last position of every sector (IndexRange::Boundary, index_range.cpp:60-66), kept in F_XE until F_LS is derived
      for (int j = tid + 1; j <= B; j += T) { fl[sector_bound(P, n, B, j) - 1] = F_XE; }
    }
    __syncthreads();
    if (!skip) {
      // ---- links, curvature (CalcCurvature, curvature.cpp:44-50: left to right from 0.0, centre weight -2P), candidates
      for (int p = tid; p < n; p += T) {
        uint32_t f = 0;
        const double r0 = rc[p].x;
        if (p + 1 < n) {
          const double r1 = rc[p + 1].x;
          if (r0 == 0.0 && r1 == 0.0) { s_cnt[0] = 1; }   // CalcRadian throws, math.cpp:40-42
          const float2 v0 = big_xy(sd, ord[p]), v1 = big_xy(sd, ord[p + 1]);
          if (link_test((double)v0.x, (double)v0.y, (double)v1.x, (double)v1.y, r0, r1, prm)) {
            f |= F_LINK;
            if (p >= P && p + 1 < n - P && !(fl[p] & F_XE)) { f |= F_LS; }
          }
        }
        double cv = 0.0;
        if (p >= P && p < n - P) {
          double sum = rc[p - P].x;
          for (int u = 1; u <= 2 * P; u++) { sum = __dadd_rn(sum, u == P ? __dmul_rn(r0, prm.center_w) : rc[p - P + u].x); }
          cv = __dmul_rn(sum, sum);
          if (cv >= prm.tau_e) { f |= F_CAND_E; }
          if (cv <= prm.tau_s) { f |= F_CAND_S; }
        }
        rc[p].y = cv;
        fl[p] = (uint8_t)f;         // (a position's sector-end mark is read by its own thread only)
      }
    }
    __syncthreads();
    if (!skip && s_cnt[0]) { skip = true; }
    if (skip) {
      for (int i = tid; i < n; i += T) {
        fl[i] = LFX_LABEL_NONE;
        if (a.sorted_src) { a.sorted_src[pos0 + i] = ord[i]; }
        if (a.curvature) { a.curvature[pos0 + i] = 0.0; }
      }
      if (tid == 0) { ring_info->status = LFX_RING_SKIPPED; ring_info->order_path = order_path; ring_info->n_edge = 0; ring_info->n_surface = 0; }
      continue;
    }

    // ---- selection. The greedy walks of label.hpp:85-94 / 124-133 over the (value, index) order are the unique solution
    //      of x_i = cand_i && no j in cover(i) with key(j) walked before key(i) and x_j (dependencies follow the strict key
    //      order, hence acyclic), reached by iterating from x = 0 in any update order - here in place, until nothing moves.
    //      `edge`: larger (curvature, index) first; surface: smaller first.
    auto sweep = [&](uint32_t cand, uint32_t x, bool edge) {
      for (;;) {
        int changed = 0;
        for (int p = tid; p < n; p += T) {
          const uint32_t f = fl[p];
          if (!(f & cand)) { continue; }
          const double cv = rc[p].y;
          bool blocked = false;
          for (int d = 1; d <= P && !blocked; d++) {          // p + d: reachable while the links hold inside the sector
            if (!(fl[p + d - 1] & F_LS)) { break; }
            if (fl[p + d] & x) { const double cj = rc[p + d].y; blocked = edge ? cj >= cv : !(cj >= cv); }   // C_d(p) = curvature(p + d) >= curvature(p), as in the on-chip kernels
          }
          for (int d = 1; d <= P && !blocked; d++) {
            if (!(fl[p - d] & F_LS)) { break; }
            if (fl[p - d] & x) { const double cj = rc[p - d].y; blocked = edge ? !(cv >= cj) : cv >= cj; }
          }
          const uint32_t fn = blocked ? (f & ~x) : (f | x);
          if (fn != f) { fl[p] = (uint8_t)fn; changed = 1; }
        }
        if (!__syncthreads_or(changed)) { break; }
      }
    };
    // cover of the picks (fill.hpp:101-117): the pick itself and up to P positions either side along unbroken links
    auto cover = [&](uint32_t x, uint32_t c) {
      for (int p = tid; p < n; p += T) {
        const uint32_t f = fl[p];
        bool in = (f & x) != 0;
        for (int d = 1; d <= P && !in; d++) {
          if (!(fl[p + d - 1] & F_LS)) { break; }
          in = (fl[p + d] & x) != 0;
        }
        for (int d = 1; d <= P && !in; d++) {
          if (p - d < 0 || !(fl[p - d] & F_LS)) { break; }
          in = (fl[p - d] & x) != 0;
        }
        if (in) { fl[p] = (uint8_t)(f | c); }
      }
      __syncthreads();
    };
    sweep(F_CAND_E, F_XE, true);
    cover(F_XE, F_CE);
    for (int p = tid; p < n; p += T) {   // still Default after the edge pass, label.hpp:125
      const uint32_t f = fl[p];
      if ((f & F_CAND_S) && (f & F_CE)) { fl[p] = (uint8_t)(f & ~F_CAND_S); }
    }
    __syncthreads();
    sweep(F_CAND_S, F_XS, false);
    cover(F_XS, F_CS);

    // ---- masks and the final label = ParallelBeam > OutOfRange > Occluded > selection (feature_extraction.cpp:133-138)
    for (int p = tid; p < n; p += T) {
      const uint32_t f = fl[p];
      const double r0 = rc[p].x;
      bool occ = false;
      {
        bool chain = true;                                      // occlusion.hpp:37-57: trigger at i = p - 1 - m
        for (int m = 0; m <= P && chain && !occ; m++) {
          const int i = p - 1 - m;
          if (i < 0) { break; }
          const bool lk = (fl[i] & F_LINK) != 0;
          if (lk && i < n - P - 1 && rc[i + 1].x > __dadd_rn(rc[i].x, prm.d)) { occ = true; }
          chain = lk;
        }
        chain = true;                                           // occlusion.hpp:59-79: trigger at i = p + 1 + m
        for (int m = 0; m <= P && chain && !occ; m++) {
          const int i = p + 1 + m;
          if (i > n - 1) { break; }
          const bool lk = (fl[i - 1] & F_LINK) != 0;
          if (lk && i >= P + 1 && rc[i - 1].x > __dadd_rn(rc[i].x, prm.d)) { occ = true; }
          chain = lk;
        }
      }
      const bool oor = !(prm.rmin <= r0 && r0 <= prm.rmax);     // out_of_range.hpp:36-48
      bool pb = false;
      if (p >= 1 && p <= n - 2) {                               // parallel_beam.hpp:36-51
        pb = ratio_test(fabs(__dsub_rn(rc[p - 1].x, r0)), r0, prm) && ratio_test(fabs(__dsub_rn(rc[p + 1].x, r0)), r0, prm);
      }
      uint32_t label = 0;                                       // Default
      if (pb) { label = 7; }
      else if (oor) { label = 5; }
      else if (occ) { label = 6; }
      else if (f & F_XE) { label = 1; }
      else if (f & F_XS) { label = 3; }
      else if (f & F_CS) { label = 4; }
      else if (f & F_CE) { label = 2; }
      a.tmp16[pos0 + p] = (uint16_t)label;
      if (a.curvature) { a.curvature[pos0 + p] = rc[p].y; }
      if (a.sorted_src) { a.sorted_src[pos0 + p] = ord[p]; }
    }
    __syncthreads();

    // ---- labels out; features staged like the on-chip ring kernel (Edge ascending from the ring start, Surface
    //      descending from the ring end: k_pack_copy moves them). Every thread owns one contiguous chunk of positions.
    const int chunk = (n + T - 1) / T;
    const int c0 = min(tid * chunk, n), c1 = min(c0 + chunk, n);
    uint32_t ne = 0, nsf = 0;
    for (int p = c0; p < c1; p++) {
      const uint32_t label = a.tmp16[pos0 + p];
      fl[p] = (uint8_t)label;
      ne += label == 1 ? 1u : 0u; nsf += label == 3 ? 1u : 0u;
    }
    s_scan[0][tid] = ne; s_scan[1][tid] = nsf;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {
      uint32_t ve = 0, vs = 0;
      if (tid >= off) { ve = s_scan[0][tid - off]; vs = s_scan[1][tid - off]; }
      __syncthreads();
      s_scan[0][tid] += ve; s_scan[1][tid] += vs;
      __syncthreads();
    }
    uint32_t re = s_scan[0][tid] - ne, rs = s_scan[1][tid] - nsf;
    for (int p = c0; p < c1; p++) {
      const uint32_t label = a.tmp16[pos0 + p];
      if (label == 1 || label == 3) {
        const uint8_t * q = sd.data + (size_t)ord[p] * sd.point_step;
        const float4 v = make_float4(*reinterpret_cast<const float *>(q + sd.off_x), *reinterpret_cast<const float *>(q + sd.off_y),
                                     *reinterpret_cast<const float *>(q + sd.off_z), 1.0f);
        if (label == 1) { a.stage[pos0 + re++] = v; } else { a.stage[pos0 + (uint32_t)(n - 1) - rs++] = v; }
      }
    }
    if (tid == T - 1) {
      ring_info->n_edge = s_scan[0][tid]; ring_info->n_surface = s_scan[1][tid];
      ring_info->status = LFX_RING_OK; ring_info->order_path = order_path;
    }
  }
}

}  // namespace lfxk
#endif  // LFX_BIG_CUH_
