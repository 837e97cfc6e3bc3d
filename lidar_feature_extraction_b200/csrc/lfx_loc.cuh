// lfx_loc.cuh — first slice of the localization consumer's hot loop: for every edge / surface
// feature of a scan, the k nearest points of the map, the line / plane model through them, and the feature's Jacobian
// block + residual of the LOAM problem. Reference: localization/include/lidar_feature_localization/edge.hpp:88-124,
// surface.hpp:106-139, localization/src/kdtree.cpp:42-55, localization/src/edge.cpp:37-83,
// rotationlib/src/jacobian/quaternion.cpp:36-52.
//
//   k_loc_knn      exact k nearest neighbours by exhaustive search: a CTA of 8 warps takes 8 queries and walks the map
//                  in tiles staged once per CTA in shared memory (SoA doubles); every lane keeps the k best of the
//                  points it saw in a sorted register list, the 32 lists are merged by k rounds of a warp argmin.
//                  Distances are nanoflann's metric_L2 over doubles (left-to-right sum of squared differences), ties
//                  go to the smaller index: the index lists equal the kd-tree's (tests: nanoflann compiled in place).
//                  Exhaustive search reads the whole map once per 8 queries (L2 resident); a grid index is the next
//                  step for maps beyond ~10^6 points.
//   k_loc_edge     per feature: mean + covariance of the neighbours, closed-form eigenvectors of the symmetric 3x3
//                  (the algorithm of Eigen's SelfAdjointEigenSolver::computeDirect), p1/p2 = mean -+ principal axis,
//                  J = [Hat(p2 - p1) dRp/dq, Hat(p2 - p1)] (3 x 7), r = (p - p1) x (p - p2).
//   k_loc_surface  per feature: plane X w = -1 by Householder QR (the algorithm of Eigen's HouseholderQR::solve),
//                  J = [u^T dRp/dq, u^T] (1 x 7) with u = w / |w|, r = (w . x + 1) / |w|.
// All arithmetic is fp64 like the reference (Eigen doubles). Eigen itself is third party and absent: its rounding is
// not reproduced bit for bit (oracle/loc_oracle.py restates the same algorithms; tolerance 1e-9 in the tests).
#ifndef LFX_LOC_CUH_
#define LFX_LOC_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace lfxk
{

constexpr int LOC_MAX_K = 16;        // neighbours per feature (the reference uses 15)
constexpr int LOC_KNN_WARPS = 8;     // queries per CTA
constexpr int LOC_TILE = 2048;       // map points per shared-memory tile

struct LocPose { double r[9]; double t[3]; double q[4]; };   // rotation matrix (row major), translation, quaternion x,y,z,w

// float x,y,z,(1) points -> SoA doubles (the reference searches a MatrixXd built from the float cloud, kdtree.hpp:72-77)
__global__ void k_loc_soa(const float4 * __restrict__ pts, uint64_t n, double * __restrict__ x, double * __restrict__ y, double * __restrict__ z)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    x[i] = (double)p.x; y[i] = (double)p.y; z[i] = (double)p.z;
  }
}

__device__ __forceinline__ void loc_transform(const LocPose & T, double px, double py, double pz, double & x, double & y, double & z)
{
  // Eigen: Isometry3d * Vector3d = linear * p + translation, row by row, left to right
  x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[0], px), __dmul_rn(T.r[1], py)), __dmul_rn(T.r[2], pz)), T.t[0]);
  y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[3], px), __dmul_rn(T.r[4], py)), __dmul_rn(T.r[5], pz)), T.t[1]);
  z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[6], px), __dmul_rn(T.r[7], py)), __dmul_rn(T.r[8], pz)), T.t[2]);
}

// (d2, idx) lexicographic order
__device__ __forceinline__ bool loc_before(double da, uint32_t ia, double db, uint32_t ib) { return da < db || (da == db && ia < ib); }

template<int K>
__global__ void __launch_bounds__(LOC_KNN_WARPS * 32)
k_loc_knn(const double * __restrict__ mx, const double * __restrict__ my, const double * __restrict__ mz, uint32_t n_map,
          const float4 * __restrict__ scan, uint32_t n_q, const LocPose T, uint32_t * __restrict__ out_idx, double * __restrict__ out_d2)
{
  __shared__ double sx[LOC_TILE], sy[LOC_TILE], sz[LOC_TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * LOC_KNN_WARPS + warp;
  double qx = 0, qy = 0, qz = 0;
  if (q < n_q) {
    const float4 p = scan[q];
    loc_transform(T, (double)p.x, (double)p.y, (double)p.z, qx, qy, qz);
  }
  double bd[K];
  uint32_t bi[K];
#pragma unroll
  for (int j = 0; j < K; j++) { bd[j] = __longlong_as_double(0x7FF0000000000000ll); bi[j] = 0xFFFFFFFFu; }
  for (uint32_t t0 = 0; t0 < n_map; t0 += LOC_TILE) {
    const uint32_t nt = min((uint32_t)LOC_TILE, n_map - t0);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) { sx[i] = mx[t0 + i]; sy[i] = my[t0 + i]; sz[i] = mz[t0 + i]; }
    __syncthreads();
    if (q < n_q) {
      for (uint32_t i = lane; i < nt; i += 32) {
        // nanoflann L2_Adaptor::evalMetric for three dimensions: result += diff * diff, dimension by dimension
        const double dx = __dsub_rn(qx, sx[i]), dy = __dsub_rn(qy, sy[i]), dz = __dsub_rn(qz, sz[i]);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        const uint32_t id = t0 + i;
        if (loc_before(d2, id, bd[K - 1], bi[K - 1])) {
          bd[K - 1] = d2; bi[K - 1] = id;
#pragma unroll
          for (int j = K - 1; j > 0; j--) {
            if (loc_before(bd[j], bi[j], bd[j - 1], bi[j - 1])) {
              const double td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
              const uint32_t ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
            }
          }
        }
      }
    }
  }
  if (q >= n_q) { return; }
  // merge: K rounds, every lane offers the head of its list
  for (int r = 0; r < K; r++) {
    double d = bd[0];
    uint32_t id = bi[0];
    int src = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xFFFFFFFFu, d, o);
      const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, id, o);
      const int os = __shfl_xor_sync(0xFFFFFFFFu, src, o);
      if (loc_before(od, oi, d, id)) { d = od; id = oi; src = os; }
    }
    if (lane == 0) { out_idx[(size_t)q * K + r] = id; out_d2[(size_t)q * K + r] = d; }
    if (lane == src) {   // the winner's list moves up
#pragma unroll
      for (int j = 0; j < K - 1; j++) { bd[j] = bd[j + 1]; bi[j] = bi[j + 1]; }
      bd[K - 1] = __longlong_as_double(0x7FF0000000000000ll); bi[K - 1] = 0xFFFFFFFFu;
    }
  }
}

// ---- exact k nearest neighbours through a uniform grid over the map (built once per lfx_loc_set_map)
//
// The map's points are bucketed into cubic cells (edge c >= 1 m, enlarged until the bounding box holds at most
// LOC_MAX_CELLS of them) and stored cell by cell (x fastest), so a row of cells along x is ONE contiguous range of points.
// A warp answers one query: it visits the cells around the query shell by shell (Chebyshev radius s = 0, 1, 2, ...),
// every lane keeping the k best of the points it saw, and stops as soon as k visited points are strictly closer than
// anything outside the visited cube can be - the distance from the query to the nearest face of the cube that still has
// cells behind it, less a safety margin that dwarfs the rounding of the cell assignment. The k best are then exactly
// those of the exhaustive search: same metric (nanoflann's L2 over doubles), same (distance, index) order.
constexpr uint32_t LOC_MAX_CELLS = 1u << 22;

struct LocGrid
{
  double x0, y0, z0;      // lower corner
  double c, inv_c;        // cell edge
  int nx, ny, nz;
  const uint32_t * cell_start;   // [nx * ny * nz + 1]
  const double * gx, * gy, * gz; // the map's points, cell by cell
  const uint32_t * gidx;         // their indices in the map
};

__device__ __forceinline__ int loc_cell1(double v, double v0, double inv_c, int n)
{
  const double t = floor((v - v0) * inv_c);
  return (int)fmin(fmax(t, 0.0), (double)(n - 1));   // (NaN -> 0)
}
__device__ __forceinline__ uint32_t loc_cell(const LocGrid & g, double x, double y, double z)
{
  const int cx = loc_cell1(x, g.x0, g.inv_c, g.nx), cy = loc_cell1(y, g.y0, g.inv_c, g.ny), cz = loc_cell1(z, g.z0, g.inv_c, g.nz);
  return ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
}

// float -> unsigned key with the same order
__device__ __forceinline__ uint32_t loc_fkey(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }

// bounding box of the finite points: keys[0..2] = min x,y,z, keys[3..5] = max x,y,z (initialised to ~0 / 0)
__global__ void __launch_bounds__(256)
k_loc_bbox(const float4 * __restrict__ pts, uint64_t n, uint32_t * __restrict__ keys)
{
  uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    const float v[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (isfinite(v[a])) { const uint32_t k = loc_fkey(v[a]); lo[a] = min(lo[a], k); hi[a] = max(hi[a], k); }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&keys[a], lo[a]); atomicMax(&keys[3 + a], hi[a]); }
  }
}

__global__ void __launch_bounds__(256)
k_loc_cell_count(const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z, uint64_t n, const LocGrid g,
                 uint32_t * __restrict__ count)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    atomicAdd(&count[loc_cell(g, x[i], y[i], z[i])], 1u);
  }
}

// one CTA: start[c] = exclusive prefix of count over the cells, start[n_cells] = total; count is cleared for the fill pass
__global__ void __launch_bounds__(1024)
k_loc_cell_scan(uint32_t * __restrict__ count, uint32_t * __restrict__ start, uint32_t n_cells)
{
  __shared__ uint32_t s_sum[1024];
  const uint32_t tid = threadIdx.x, per = (n_cells + 1023u) / 1024u;
  const uint32_t b0 = min(tid * per, n_cells), b1 = min(b0 + per, n_cells);
  uint32_t sum = 0;
  for (uint32_t i = b0; i < b1; i++) { sum += count[i]; }
  s_sum[tid] = sum;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    const uint32_t v = tid >= off ? s_sum[tid - off] : 0u;
    __syncthreads();
    s_sum[tid] += v;
    __syncthreads();
  }
  uint32_t run = s_sum[tid] - sum;
  for (uint32_t i = b0; i < b1; i++) { const uint32_t v = count[i]; start[i] = run; run += v; count[i] = 0; }
  if (tid == 1023) { start[n_cells] = s_sum[1023]; }
}

__global__ void __launch_bounds__(256)
k_loc_cell_fill(const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z, uint64_t n, const LocGrid g,
                uint32_t * __restrict__ fill, double * __restrict__ gx, double * __restrict__ gy, double * __restrict__ gz, uint32_t * __restrict__ gidx)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double px = x[i], py = y[i], pz = z[i];
    const uint32_t c = loc_cell(g, px, py, pz);
    const uint32_t at = g.cell_start[c] + atomicAdd(&fill[c], 1u);
    gx[at] = px; gy[at] = py; gz[at] = pz; gidx[at] = (uint32_t)i;
  }
}

template<int K>
__global__ void __launch_bounds__(LOC_KNN_WARPS * 32)
k_loc_knn_grid(const LocGrid g, const float4 * __restrict__ scan, uint32_t n_q, const LocPose T, uint32_t * __restrict__ out_idx,
               double * __restrict__ out_d2)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x * LOC_KNN_WARPS + warp;
  if (q >= n_q) { return; }
  double qx, qy, qz;
  {
    const float4 p = scan[q];
    loc_transform(T, (double)p.x, (double)p.y, (double)p.z, qx, qy, qz);
  }
  const int cx = loc_cell1(qx, g.x0, g.inv_c, g.nx), cy = loc_cell1(qy, g.y0, g.inv_c, g.ny), cz = loc_cell1(qz, g.z0, g.inv_c, g.nz);
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  double bd[K];
  uint32_t bi[K];
#pragma unroll
  for (int j = 0; j < K; j++) { bd[j] = inf; bi[j] = 0xFFFFFFFFu; }
  // all points of the cells [xa, xb] of row (y, z): one contiguous range
  auto visit = [&](int xa, int xb, int y, int z) {
    xa = max(xa, 0); xb = min(xb, g.nx - 1);
    if (xa > xb) { return; }
    const uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
    const uint32_t a = g.cell_start[row + (uint32_t)xa], b = g.cell_start[row + (uint32_t)xb + 1u];
    for (uint32_t i = a + (uint32_t)lane; i < b; i += 32) {
      // nanoflann L2_Adaptor::evalMetric for three dimensions: result += diff * diff, dimension by dimension
      const double dx = __dsub_rn(qx, g.gx[i]), dy = __dsub_rn(qy, g.gy[i]), dz = __dsub_rn(qz, g.gz[i]);
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      const uint32_t id = g.gidx[i];
      if (loc_before(d2, id, bd[K - 1], bi[K - 1])) {
        bd[K - 1] = d2; bi[K - 1] = id;
#pragma unroll
        for (int j = K - 1; j > 0; j--) {
          if (loc_before(bd[j], bi[j], bd[j - 1], bi[j - 1])) {
            const double td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
            const uint32_t ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
          }
        }
      }
    }
  };
  for (int s = 0;; s++) {
    for (int dz = -s; dz <= s; dz++) {
      const int z = cz + dz;
      if (z < 0 || z >= g.nz) { continue; }
      for (int dy = -s; dy <= s; dy++) {
        const int y = cy + dy;
        if (y < 0 || y >= g.ny) { continue; }
        if (s == 0 || dz == -s || dz == s || dy == -s || dy == s) { visit(cx - s, cx + s, y, z); }   // a face of the shell
        else { visit(cx - s, cx - s, y, z); visit(cx + s, cx + s, y, z); }                           // its two end cells
      }
    }
    // nothing outside the cube [c - s, c + s]^3 is closer than the nearest of its faces that has cells behind it
    double bound = inf;
    if (cx - s > 0) { bound = fmin(bound, qx - (g.x0 + (double)(cx - s) * g.c)); }
    if (cx + s < g.nx - 1) { bound = fmin(bound, (g.x0 + (double)(cx + s + 1) * g.c) - qx); }
    if (cy - s > 0) { bound = fmin(bound, qy - (g.y0 + (double)(cy - s) * g.c)); }
    if (cy + s < g.ny - 1) { bound = fmin(bound, (g.y0 + (double)(cy + s + 1) * g.c) - qy); }
    if (cz - s > 0) { bound = fmin(bound, qz - (g.z0 + (double)(cz - s) * g.c)); }
    if (cz + s < g.nz - 1) { bound = fmin(bound, (g.z0 + (double)(cz + s + 1) * g.c) - qz); }
    if (bound == inf) { break; }                       // the cube covers the whole grid
    bound = bound - 1.0e-6 * g.c;                      // (cell assignment rounds at ~1e-16 of the coordinates)
    int closer = 0;
    if (bound > 0.0) {
      const double b2 = bound * bound;
#pragma unroll
      for (int j = 0; j < K; j++) { closer += bd[j] < b2 ? 1 : 0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { closer += __shfl_xor_sync(0xFFFFFFFFu, closer, o); }
    if (closer >= K) { break; }
  }
  // merge: K rounds, every lane offers the head of its list
  for (int r = 0; r < K; r++) {
    double d = bd[0];
    uint32_t id = bi[0];
    int src = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xFFFFFFFFu, d, o);
      const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, id, o);
      const int os = __shfl_xor_sync(0xFFFFFFFFu, src, o);
      if (loc_before(od, oi, d, id)) { d = od; id = oi; src = os; }
    }
    if (lane == 0) { out_idx[(size_t)q * K + r] = id; out_d2[(size_t)q * K + r] = d; }
    if (lane == src) {   // the winner's list moves up
#pragma unroll
      for (int j = 0; j < K - 1; j++) { bd[j] = bd[j + 1]; bi[j] = bi[j + 1]; }
      bd[K - 1] = inf; bi[K - 1] = 0xFFFFFFFFu;
    }
  }
}

// ---- small dense pieces (one thread per feature)

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 cross(const V3 & a, const V3 & b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double dot(const V3 & a, const V3 & b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// rotationlib::DRpDq (quaternion.cpp:36-52), J[3][4], columns (w, x, y, z)
__device__ __forceinline__ void loc_drpdq(const double * q /* x,y,z,w */, const V3 & p, double (&J)[3][4])
{
  const V3 v = v3(q[0], q[1], q[2]);
  const double w = q[3];
  const V3 c = cross(v, p);
  const double vp = dot(v, p);
  const double pv[3] = {p.x, p.y, p.z}, vv[3] = {v.x, v.y, v.z};
  const double K[3][3] = {{0.0, -p.z, p.y}, {p.z, 0.0, -p.x}, {-p.y, p.x, 0.0}};
  J[0][0] = 2.0 * (w * p.x + c.x); J[1][0] = 2.0 * (w * p.y + c.y); J[2][0] = 2.0 * (w * p.z + c.z);
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      J[i][1 + j] = 2.0 * ((i == j ? vp : 0.0) + vv[i] * pv[j] - pv[i] * vv[j] - w * K[i][j]);
    }
  }
}

// internal::direct_selfadjoint_eigenvalues<SolverType, 3, false>: roots of the characteristic polynomial, ascending
__device__ __forceinline__ void loc_roots(const double (&m)[3][3], double (&roots)[3])
{
  const double s_inv3 = 1.0 / 3.0, s_sqrt3 = sqrt(3.0);
  const double c0 = m[0][0] * m[1][1] * m[2][2] + 2.0 * m[1][0] * m[2][0] * m[2][1] - m[0][0] * m[2][1] * m[2][1] -
                    m[1][1] * m[2][0] * m[2][0] - m[2][2] * m[1][0] * m[1][0];
  const double c1 = m[0][0] * m[1][1] - m[1][0] * m[1][0] + m[0][0] * m[2][2] - m[2][0] * m[2][0] + m[1][1] * m[2][2] - m[2][1] * m[2][1];
  const double c2 = m[0][0] + m[1][1] + m[2][2];
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
  a_over_3 = fmax(a_over_3, 0.0);
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double qq = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
  qq = fmax(qq, 0.0);
  const double rho = sqrt(a_over_3);
  const double theta = atan2(sqrt(qq), half_b) * s_inv3;
  const double ct = cos(theta), st = sin(theta);
  roots[0] = c2_over_3 - rho * (ct + s_sqrt3 * st);
  roots[1] = c2_over_3 - rho * (ct - s_sqrt3 * st);
  roots[2] = c2_over_3 + 2.0 * rho * ct;
}

// extract_kernel: null-space direction of a rank-2 symmetric 3x3 from the cross products of its columns
__device__ __forceinline__ V3 loc_extract_kernel(const double (&m)[3][3])
{
  int i0 = 0;
  if (fabs(m[1][1]) > fabs(m[i0][i0])) { i0 = 1; }
  if (fabs(m[2][2]) > fabs(m[i0][i0])) { i0 = 2; }
  const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  const V3 rep = v3(m[0][i0], m[1][i0], m[2][i0]);
  const V3 c0 = cross(rep, v3(m[0][i1], m[1][i1], m[2][i1]));
  const V3 c1 = cross(rep, v3(m[0][i2], m[1][i2], m[2][i2]));
  const double n0 = dot(c0, c0), n1 = dot(c1, c1);
  if (n0 > n1) { const double s = 1.0 / sqrt(n0); return v3(c0.x * s, c0.y * s, c0.z * s); }
  const double s = 1.0 / sqrt(n1);
  return v3(c1.x * s, c1.y * s, c1.z * s);
}

// column 2 (largest eigenvalue) of SelfAdjointEigenSolver<Matrix3d>::computeDirect(C).eigenvectors()
__device__ __forceinline__ V3 loc_principal_axis(const double (&C)[3][3])
{
  const double eps = 2.220446049250313e-16;
  const double shift = (C[0][0] + C[1][1] + C[2][2]) / 3.0;
  double m[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) { m[i][j] = (i >= j ? C[i][j] : C[j][i]) - (i == j ? shift : 0.0); }   // lower triangle
  }
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) { scale = fmax(scale, fabs(m[i][j])); }
  }
  if (scale > 0.0) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) { m[i][j] /= scale; }
    }
  }
  double ev[3];
  loc_roots(m, ev);
  if (!(ev[2] - ev[0] > eps)) { return v3(0.0, 0.0, 1.0); }   // eigenvectors = identity
  // eigenvector of the largest eigenvalue: the null space of (scaled matrix - ev[2] I). In the solver it is either
  // the "most distinct" one (computed first) or the second one (computed the same way unless the other two
  // eigenvalues coincide numerically, in which case it is rebuilt from the first vector's representative).
  const double d0 = ev[2] - ev[1], d1 = ev[1] - ev[0];
  const bool k_is_2 = d0 > d1;
  double tmp[3][3];
  if (k_is_2 || !(d0 <= 2.0 * eps * d1)) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) { tmp[i][j] = m[i][j] - (i == j ? ev[2] : 0.0); }
    }
    return loc_extract_kernel(tmp);
  }
  // k = 0, l = 2 and d0 (= ev[2] - ev[1]) negligible against d1: col(2) = rep - (col(0) . rep) rep, normalised, where
  // rep is the representative column extract_kernel chose for (scaled matrix - ev[0] I)
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) { tmp[i][j] = m[i][j] - (i == j ? ev[0] : 0.0); }
  }
  const V3 e0 = loc_extract_kernel(tmp);
  int i0 = 0;
  if (fabs(tmp[1][1]) > fabs(tmp[i0][i0])) { i0 = 1; }
  if (fabs(tmp[2][2]) > fabs(tmp[i0][i0])) { i0 = 2; }
  const V3 rep = v3(tmp[0][i0], tmp[1][i0], tmp[2][i0]);
  const double a = dot(e0, rep);
  V3 v = v3(rep.x - a * rep.x, rep.y - a * rep.y, rep.z - a * rep.z);
  const double s = 1.0 / sqrt(dot(v, v));
  return v3(v.x * s, v.y * s, v.z * s);
}

struct LocArgs
{
  const double * mx, * my, * mz;
  const float4 * scan;
  const uint32_t * nbr;   // [n][k]
  uint32_t n;
  int k;
  LocPose T;
  double * J;             // edge: [n][3][7]; surface: [n][7]
  double * r;             // edge: [n][3];    surface: [n]
};

__global__ void __launch_bounds__(128)
k_loc_edge(const LocArgs a)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) { return; }
  const uint32_t * nb = a.nbr + (size_t)i * a.k;
  // CalcMeanAndCovariance, edge.cpp:42-48: mean = column means, covariance = D^T D / n
  double sx = 0, sy = 0, sz = 0;
  for (int j = 0; j < a.k; j++) { sx += a.mx[nb[j]]; sy += a.my[nb[j]]; sz += a.mz[nb[j]]; }
  const double inv = (double)a.k;
  const V3 mean = v3(sx / inv, sy / inv, sz / inv);
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < a.k; j++) {
    const double d[3] = {a.mx[nb[j]] - mean.x, a.my[nb[j]] - mean.y, a.mz[nb[j]] - mean.z};
#pragma unroll
    for (int u = 0; u < 3; u++) {
#pragma unroll
      for (int v = 0; v < 3; v++) { C[u][v] += d[u] * d[v]; }
    }
  }
#pragma unroll
  for (int u = 0; u < 3; u++) {
#pragma unroll
    for (int v = 0; v < 3; v++) { C[u][v] /= inv; }
  }
  const V3 pr = loc_principal_axis(C);
  const float4 sp = a.scan[i];
  const V3 p0 = v3((double)sp.x, (double)sp.y, (double)sp.z);
  const V3 p1 = v3(mean.x - pr.x, mean.y - pr.y, mean.z - pr.z), p2 = v3(mean.x + pr.x, mean.y + pr.y, mean.z + pr.z);
  // MakeEdgeJacobianRow, edge.cpp:64-73
  const V3 dd = v3(p2.x - p1.x, p2.y - p1.y, p2.z - p1.z);
  const double K[3][3] = {{0.0, -dd.z, dd.y}, {dd.z, 0.0, -dd.x}, {-dd.y, dd.x, 0.0}};
  double D[3][4];
  loc_drpdq(a.T.q, p0, D);
  double * J = a.J + (size_t)i * 21;
#pragma unroll
  for (int u = 0; u < 3; u++) {
#pragma unroll
    for (int v = 0; v < 4; v++) { J[u * 7 + v] = K[u][0] * D[0][v] + K[u][1] * D[1][v] + K[u][2] * D[2][v]; }
#pragma unroll
    for (int v = 0; v < 3; v++) { J[u * 7 + 4 + v] = K[u][v]; }
  }
  // MakeEdgeResidual, edge.cpp:75-83
  double px, py, pz;
  loc_transform(a.T, p0.x, p0.y, p0.z, px, py, pz);
  const V3 r = cross(v3(px - p1.x, py - p1.y, pz - p1.z), v3(px - p2.x, py - p2.y, pz - p2.z));
  a.r[(size_t)i * 3] = r.x; a.r[(size_t)i * 3 + 1] = r.y; a.r[(size_t)i * 3 + 2] = r.z;
}

__global__ void __launch_bounds__(128)
k_loc_surface(const LocArgs a)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) { return; }
  const uint32_t * nb = a.nbr + (size_t)i * a.k;
  // EstimatePlaneCoefficients, surface.hpp:78-82: X w = -1 in the least-squares sense by Householder QR (math.hpp:36-40)
  double A[LOC_MAX_K][3], c[LOC_MAX_K];
  const int m = a.k;
  for (int j = 0; j < m; j++) { A[j][0] = a.mx[nb[j]]; A[j][1] = a.my[nb[j]]; A[j][2] = a.mz[nb[j]]; c[j] = -1.0; }
  double diag[3];
#pragma unroll
  for (int kk = 0; kk < 3; kk++) {
    double tail2 = 0.0;
    for (int j = kk + 1; j < m; j++) { tail2 += A[j][kk] * A[j][kk]; }
    const double c0 = A[kk][kk];
    double tau = 0.0, beta = c0;
    if (tail2 > 2.2250738585072014e-308) {   // makeHouseholder
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) { beta = -beta; }
      for (int j = kk + 1; j < m; j++) { A[j][kk] /= (c0 - beta); }
      tau = (beta - c0) / beta;
    } else {
      for (int j = kk + 1; j < m; j++) { A[j][kk] = 0.0; }
    }
    // applyHouseholderOnTheLeft on the remaining columns and on the right-hand side
    for (int col = kk + 1; col <= 3; col++) {
      double s = col < 3 ? A[kk][col] : c[kk];
      for (int j = kk + 1; j < m; j++) { s += A[j][kk] * (col < 3 ? A[j][col] : c[j]); }
      s *= tau;
      if (col < 3) { A[kk][col] -= s; } else { c[kk] -= s; }
      for (int j = kk + 1; j < m; j++) {
        if (col < 3) { A[j][col] -= s * A[j][kk]; } else { c[j] -= s * A[j][kk]; }
      }
    }
    diag[kk] = beta;
  }
  double w[3];
  w[2] = c[2] / diag[2];
  w[1] = (c[1] - A[1][2] * w[2]) / diag[1];
  w[0] = (c[0] - (A[0][1] * w[1] + A[0][2] * w[2])) / diag[0];
  const double norm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double u[3] = {w[0] / norm, w[1] / norm, w[2] / norm};
  const float4 sp = a.scan[i];
  const V3 p = v3((double)sp.x, (double)sp.y, (double)sp.z);
  double D[3][4];
  loc_drpdq(a.T.q, p, D);
  double * J = a.J + (size_t)i * 7;   // MakeJacobianRow, surface.hpp:84-92
#pragma unroll
  for (int v = 0; v < 4; v++) { J[v] = u[0] * D[0][v] + u[1] * D[1][v] + u[2] * D[2][v]; }
  J[4] = u[0]; J[5] = u[1]; J[6] = u[2];
  double x, y, z;
  loc_transform(a.T, p.x, p.y, p.z, x, y, z);
  a.r[i] = (w[0] * x + w[1] * y + w[2] * z + 1.0) / norm;   // SignedPointPlaneDistance, surface.hpp:46-50
}

}  // namespace lfxk
#endif  // LFX_LOC_CUH_
