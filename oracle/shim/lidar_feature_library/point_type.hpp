// Shim: shadows lib/include/lidar_feature_library/point_type.hpp (which needs PCL's
// registration macros) with a layout-identical PointXYZIR: 16-byte aligned xyz + 1.0f,
// then intensity (f32) and ring (u16); sizeof == 32. Test infrastructure only.
#ifndef LFX_SHIM_POINT_TYPE_HPP_
#define LFX_SHIM_POINT_TYPE_HPP_
#include <cstdint>
struct alignas(16) PointXYZIR
{
  float x, y, z, w;
  float intensity;
  std::uint16_t ring;
  PointXYZIR(float _x, float _y, float _z, float _intensity, std::uint16_t _ring)
  : x(_x), y(_y), z(_z), w(1.0f), intensity(_intensity), ring(_ring) {}
  PointXYZIR() : PointXYZIR(0.f, 0.f, 0.f, 0.f, 0) {}
};
static_assert(sizeof(PointXYZIR) == 32, "PointXYZIR must be 32 bytes");
#endif
