// Shim: sensor_msgs::msg::PointField with the setter chain test_ring.cpp uses. Test infrastructure only.
#ifndef LFX_SHIM_POINT_FIELD_HPP_
#define LFX_SHIM_POINT_FIELD_HPP_
#include <cstdint>
#include <string>
namespace sensor_msgs
{
namespace msg
{
struct PointField
{
  std::string name;
  std::uint32_t offset = 0;
  std::uint8_t datatype = 0;
  std::uint32_t count = 0;
  PointField & set__name(const std::string & v) {name = v; return *this;}
  PointField & set__offset(std::uint32_t v) {offset = v; return *this;}
  PointField & set__datatype(std::uint8_t v) {datatype = v; return *this;}
  PointField & set__count(std::uint32_t v) {count = v; return *this;}
};
}  // namespace msg
}  // namespace sensor_msgs
#endif
