// lfx_color.cuh — the colored_scan debug cloud on the device.
//
// Replaces ColorPointsByLabel / MakeXYZRGB (extraction/include/lidar_feature_extraction/color_points.hpp:46-74)
// and LabelToColor (extraction/src/color_points.cpp:39-68) as used by the node (feature_extraction.cpp:153,161):
// every point of every ring that contributes (not sparse, not skipped), rings ascending, ring-sorted order,
// as a 32-byte pcl::PointXYZRGB: x,y,z, 1.0f, {b,g,r,a=255}, 12 zero bytes - the bytes pcl::toROSMsg puts on the
// wire for the fields x,y,z (FLOAT32 @0,4,8) and rgb (FLOAT32 @16), point_step 32.
// Byte work bound by HBM: 1 label byte + 4 index bytes + one 32-byte sector of the source point read, 32 bytes
// written per point.
#ifndef LFX_COLOR_CUH_
#define LFX_COLOR_CUH_

#include "lfx_kernels.cuh"

namespace lfxk
{

struct ColorArgs
{
  const ScanDesc * scans;
  const lfx_ring_info * rings;
  const uint8_t * labels;
  const uint32_t * sorted_src;
  uint4 * out;            // [total_points][2]: scan s starts at point_base[s]
  uint32_t * counts;      // [n_scans]
  int max_rings;
};

constexpr int COLOR_THREADS = 256;

// grid (ring id, scan)
__global__ void __launch_bounds__(COLOR_THREADS)
k_color_scan(const ColorArgs a)
{
  const int r = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const lfx_ring_info * rings = a.rings + (size_t)s * a.max_rings;
  const lfx_ring_info ri = rings[r];
  const bool mine = ri.count > 0 && ri.status == LFX_RING_OK;
  const bool totals = r == 0;   // the first CTA of a scan also reports the scan's width
  if (!mine && !totals) { return; }
  // points of contributing rings with a smaller id come first (totals: of all rings)
  uint32_t before = 0;
  const int upto = totals && !mine ? a.max_rings : r;
  uint32_t all = 0;
  for (int q = tid; q < a.max_rings; q += COLOR_THREADS) {
    const lfx_ring_info o = rings[q];
    const uint32_t n = o.status == LFX_RING_OK ? o.count : 0u;
    if (q < upto) { before += n; }
    all += n;
  }
  unsigned long long both = (unsigned long long)before | ((unsigned long long)all << 32);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { both += __shfl_xor_sync(0xFFFFFFFFu, both, o); }
  __shared__ unsigned long long s_both[COLOR_THREADS / 32];
  if (lane == 0) { s_both[warp] = both; }
  __syncthreads();
  both = 0;
#pragma unroll
  for (int w = 0; w < COLOR_THREADS / 32; w++) { both += s_both[w]; }
  before = (uint32_t)both;
  if (totals && tid == 0) { a.counts[s] = (uint32_t)(both >> 32); }
  if (!mine) { return; }
  const ScanDesc sd = a.scans[s];
  const uint64_t pos0 = sd.point_base + ri.offset;
  uint4 * out = a.out + 2 * (sd.point_base + before);
  // LabelToColor, color_points.cpp:39-68, packed as pcl::PointXYZRGB's {b, g, r, a = 255}
  const uint32_t bgra[8] = {0xFFFFFFFFu, 0xFFFF0000u, 0xFFFF3F00u, 0xFFFF0000u, 0xFFFF3F00u, 0xFF7F7F7Fu, 0xFFFF00FFu, 0xFF00FF00u};
  for (uint32_t p = tid; p < ri.count; p += COLOR_THREADS) {
    const uint32_t src = a.sorted_src[pos0 + p];
    const uint32_t label = a.labels[pos0 + p];
    const uint8_t * pt = sd.data + (size_t)src * sd.point_step;
    float x, y, z;
    if (sd.vec_ok) {
      const float4 v = *reinterpret_cast<const float4 *>(pt + sd.off_x);
      x = v.x; y = v.y; z = v.z;
    } else {
      x = *reinterpret_cast<const float *>(pt + sd.off_x);
      y = *reinterpret_cast<const float *>(pt + sd.off_y);
      z = *reinterpret_cast<const float *>(pt + sd.off_z);
    }
    __stcs(out + 2 * (size_t)p, make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), 0x3F800000u));
    __stcs(out + 2 * (size_t)p + 1, make_uint4(bgra[label & 7u], 0u, 0u, 0u));
  }
}

}  // namespace lfxk
#endif  // LFX_COLOR_CUH_
