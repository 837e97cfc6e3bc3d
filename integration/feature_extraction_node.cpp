// feature_extraction_node.cpp — what extraction/app/feature_extraction.cpp becomes when the per-scan work of
// its callback (lines 110-157 + 163-164) is replaced by the B200 library. SOURCE ONLY: this image has no
// ROS 2 / PCL, so the file is not compiled here; everything ROS-facing (topics, QoS, parameter names, frame
// id, stamps, shutdown behaviour) is unchanged from the reference, cited line by line.
#include <memory>
#include <string>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/msg/point_cloud2.hpp>

#include "lfx.hpp"

static lfx::HyperParameters DeclareParameters(rclcpp::Node & node)
{
  // hyper_parameter.hpp:35-43 — same names, same defaults (taken from lfx_default_params)
  lfx::HyperParameters p;
  p.padding = node.declare_parameter("convolution_padding", p.padding);
  p.neighbor_degree_threshold = node.declare_parameter("neighbor_degree_threshold", p.neighbor_degree_threshold);
  p.distance_diff_threshold = node.declare_parameter("distance_diff_threshold", p.distance_diff_threshold);
  p.parallel_beam_min_range_ratio = node.declare_parameter("parallel_beam_min_range_ratio", p.parallel_beam_min_range_ratio);
  p.edge_threshold = node.declare_parameter("edge_threshold", p.edge_threshold);
  p.surface_threshold = node.declare_parameter("surface_threshold", p.surface_threshold);
  p.min_range = node.declare_parameter("min_range", p.min_range);
  p.max_range = node.declare_parameter("max_range", p.max_range);
  p.n_blocks = node.declare_parameter("n_blocks", p.n_blocks);
  return p;  // lfx_create re-checks positivity (hyper_parameter.hpp:45-53) and returns LFX_E_BAD_PARAM
}

// any of the three output topics from the library's byte buffer and lfx_topic_layout (ros_msg.hpp:53-71)
static sensor_msgs::msg::PointCloud2 MakeCloud(int topic, const uint8_t * data, uint32_t n, const std_msgs::msg::Header & header)
{
  const lfx::TopicLayout layout(topic);
  sensor_msgs::msg::PointCloud2 msg;
  msg.header = header;
  msg.height = 1;
  msg.width = n;
  msg.is_dense = true;
  msg.is_bigendian = false;
  msg.point_step = layout.point_step;
  msg.row_step = layout.point_step * n;
  for (const auto & lf : layout.fields) {
    sensor_msgs::msg::PointField f;
    f.name = lf.name; f.offset = lf.offset; f.datatype = lf.datatype; f.count = 1;
    msg.fields.push_back(f);
  }
  msg.data.assign(data, data + static_cast<size_t>(layout.point_step) * n);
  return msg;
}

static sensor_msgs::msg::PointCloud2 MakeXYZCloud(const float * xyzw, uint32_t n, const std_msgs::msg::Header & header)
{
  // what pcl::toROSMsg emits for pcl::PointXYZ (ros_msg.hpp:53-71): x,y,z FLOAT32 at 0/4/8, point_step 16
  sensor_msgs::msg::PointCloud2 msg;
  msg.header = header;
  msg.height = 1;
  msg.width = n;
  msg.is_dense = true;
  msg.is_bigendian = false;
  msg.point_step = 16;
  msg.row_step = 16 * n;
  const char * names[3] = {"x", "y", "z"};
  for (uint32_t k = 0; k < 3; k++) {
    sensor_msgs::msg::PointField f;
    f.name = names[k]; f.offset = 4 * k; f.datatype = sensor_msgs::msg::PointField::FLOAT32; f.count = 1;
    msg.fields.push_back(f);
  }
  msg.data.assign(reinterpret_cast<const uint8_t *>(xyzw), reinterpret_cast<const uint8_t *>(xyzw) + 16 * static_cast<size_t>(n));
  return msg;
}

class FeatureExtractionNode : public rclcpp::Node
{
public:
  FeatureExtractionNode()
  : Node("lidar_feature_extraction"), fe_(DeclareParameters(*this), DebugOptions())
  {
    // feature_extraction.cpp:73-82 — same topics and QoS
    const auto qos = rclcpp::SensorDataQoS().reliable().durability_volatile();
    sub_ = create_subscription<sensor_msgs::msg::PointCloud2>(
      "points_raw", qos, std::bind(&FeatureExtractionNode::Callback, this, std::placeholders::_1));
    edge_pub_ = create_publisher<sensor_msgs::msg::PointCloud2>("scan_edge", qos);
    surface_pub_ = create_publisher<sensor_msgs::msg::PointCloud2>("scan_surface", qos);
    colored_pub_ = create_publisher<sensor_msgs::msg::PointCloud2>("colored_scan", 1);   // :77-78
  }

private:
  static lfx_options DebugOptions()
  {
    lfx_options opt{};
    opt.want_sorted_src = 1;   // keeps the ring-sorted -> source index map that colored_scan needs
    return opt;
  }

  void Callback(const sensor_msgs::msg::PointCloud2::ConstSharedPtr msg)
  {
    std::vector<lfx::PointField> fields;
    for (const auto & f : msg->fields) { fields.push_back({f.name, f.offset, f.datatype}); }
    lfx_scan_output out;
    try {
      const lfx_cloud_view view = lfx::MakeView(msg->data.data(), msg->width * msg->height, msg->point_step, fields, msg->is_dense, false, msg->data.size());
      out = fe_.Extract(view);   // replaces feature_extraction.cpp:110-157
    } catch (const lfx::Error & e) {
      // LFX_E_NOT_DENSE / LFX_E_NO_RING: feature_extraction.cpp:96-108 logs and shuts down
      RCLCPP_ERROR(get_logger(), "%s", e.what());
      rclcpp::shutdown();
      return;
    }
    // feature_extraction.cpp:154-156: one warning per ring whose per-ring code throws in the reference (too few points
    // for the convolution / the sectors, two adjacent points on the sensor axis); such a ring contributes nothing
    for (const int ring : fe_.SkippedRings()) {
      RCLCPP_WARN(get_logger(), "ring %d was skipped (the reference throws std::invalid_argument for it)", ring);
    }
    std_msgs::msg::Header header = msg->header;       // stamp = input stamp (consumers sync on it, subscriber.hpp:72-75)
    header.frame_id = "lidar_feature_base_link";      // feature_extraction.cpp:159
    edge_pub_->publish(MakeXYZCloud(out.edge_xyz, out.n_edge, header));           // :163-170
    surface_pub_->publish(MakeXYZCloud(out.surface_xyz, out.n_surface, header));
    // colored_scan (depth-1 debug topic, :77-78,153,161,168): built on the device when the handle was created with
    // lfx_options.want_sorted_src; the bytes are PointCloud2.data of the pcl::PointXYZRGB message
    // (published on every callback like the reference does, :168, whether or not anybody listens)
    const std::vector<uint8_t> colored = fe_.ColoredScan();
    colored_pub_->publish(MakeCloud(LFX_TOPIC_COLORED_SCAN, colored.data(), static_cast<uint32_t>(colored.size() / 32), header));
  }

  lfx::FeatureExtraction fe_;
  rclcpp::Subscription<sensor_msgs::msg::PointCloud2>::SharedPtr sub_;
  rclcpp::Publisher<sensor_msgs::msg::PointCloud2>::SharedPtr edge_pub_, surface_pub_, colored_pub_;
};

int main(int argc, char * argv[])
{
  rclcpp::init(argc, argv);
  rclcpp::spin(std::make_shared<FeatureExtractionNode>());  // feature_extraction.cpp:182-188
  rclcpp::shutdown();
  return 0;
}
