// Shim: sensor_msgs::msg::PointCloud2 (fields + payload only). Test infrastructure only.
#ifndef LFX_SHIM_POINT_CLOUD2_HPP_
#define LFX_SHIM_POINT_CLOUD2_HPP_
#include <cstdint>
#include <vector>
#include "sensor_msgs/msg/point_field.hpp"
namespace sensor_msgs
{
namespace msg
{
struct PointCloud2
{
  std::uint32_t height = 1;
  std::uint32_t width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  std::uint32_t point_step = 0;
  std::uint32_t row_step = 0;
  std::vector<std::uint8_t> data;
  bool is_dense = true;
};
}  // namespace msg
}  // namespace sensor_msgs
#endif
