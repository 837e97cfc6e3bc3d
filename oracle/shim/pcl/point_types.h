// Shim: minimal stand-in for <pcl/point_types.h> (test infrastructure).
#ifndef LFX_SHIM_PCL_POINT_TYPES_H_
#define LFX_SHIM_PCL_POINT_TYPES_H_
#include <cstdint>
namespace pcl
{
struct alignas(16) PointXYZ
{
  float x, y, z, w;
  PointXYZ() : x(0.f), y(0.f), z(0.f), w(1.f) {}
  PointXYZ(float _x, float _y, float _z) : x(_x), y(_y), z(_z), w(1.f) {}
};
struct alignas(16) PointXYZRGB
{
  float x, y, z, w;
  std::uint8_t b, g, r, a;
  float pad[3];
  PointXYZRGB() : x(0.f), y(0.f), z(0.f), w(1.f), b(0), g(0), r(0), a(255), pad{0.f, 0.f, 0.f} {}
};
}  // namespace pcl
#endif
