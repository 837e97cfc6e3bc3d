// lfx_map.cuh — mapping accumulate on the device: Map::TransformAdd
// (mapping/include/lidar_feature_mapping/map.hpp:68-74) for the frames MapBuilder::Callback (:104-127) accepts.
// The pose gate is a sequential scan over the frames and runs on the host (lfx_api.cu: map_gate); here every accepted
// frame's scan_edge cloud is moved by its pose, in double like pcl::transformPointCloud with an Affine3d
// (out = (float)(((m0 x + m1 y) + m2 z) + m3), uncontracted), and appended to the map in frame order.
// Byte work: 16 B read + 16 B written per edge point.
#ifndef LFX_MAP_CUH_
#define LFX_MAP_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace lfxk
{

struct MapFrame
{
  double m[12];        // row-major 3x4 [R | t]
  uint64_t dst;        // first map position of this frame's points
  uint32_t src, n;     // the frame's points in the batch's concatenated edge cloud
};

constexpr int MAP_THREADS = 128;

__global__ void __launch_bounds__(MAP_THREADS)
k_map_transform_add(const MapFrame * frames, const float4 * edge, float4 * map)
{
  __shared__ MapFrame f;
  if (threadIdx.x < sizeof(MapFrame) / 8) { reinterpret_cast<uint64_t *>(&f)[threadIdx.x] = reinterpret_cast<const uint64_t *>(frames + blockIdx.x)[threadIdx.x]; }
  __syncthreads();
  const float4 * src = edge + f.src;
  float4 * dst = map + f.dst;
  for (uint32_t i = blockIdx.y * MAP_THREADS + threadIdx.x; i < f.n; i += gridDim.y * MAP_THREADS) {
    const float4 p = src[i];
    const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
    float4 o;
    o.x = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(f.m[0], x), __dmul_rn(f.m[1], y)), __dmul_rn(f.m[2], z)), f.m[3]));
    o.y = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(f.m[4], x), __dmul_rn(f.m[5], y)), __dmul_rn(f.m[6], z)), f.m[7]));
    o.z = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(f.m[8], x), __dmul_rn(f.m[9], y)), __dmul_rn(f.m[10], z)), f.m[11]));
    o.w = 1.0f;
    dst[i] = o;
  }
}

}  // namespace lfxk
#endif  // LFX_MAP_CUH_
