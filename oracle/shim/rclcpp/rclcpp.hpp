// Shim: label.hpp includes rclcpp but never uses it. Test infrastructure only.
#ifndef LFX_SHIM_RCLCPP_HPP_
#define LFX_SHIM_RCLCPP_HPP_
#endif
