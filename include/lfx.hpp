// lfx.hpp — header-only C++17 host mirror of the reference extraction node's state over the C ABI (lfx.h).
//
// Mirrors, for the one path this library replaces:
//   lfx::HyperParameters   extraction/include/lidar_feature_extraction/hyper_parameter.hpp:32-65
//   lfx::FeatureExtraction the per-scan work of FeatureExtraction::Callback, extraction/app/feature_extraction.cpp:92-171
//   lfx::PointTypeConverter the upstream converter node, point_type_converter/point_type_converter/convert.py:171-212
//   lfx::TopicLayout / FeatureExtraction::ColoredScan  ToRosMsg<T> (ros_msg.hpp:53-71), ColorPointsByLabel (color_points.hpp:60-74)
// No CUDA or ROS headers are needed to include this file; link with -llfx.
#ifndef LFX_HPP_
#define LFX_HPP_

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "lfx.h"

namespace lfx
{

struct Error : std::runtime_error
{
  int code;
  Error(int c, const std::string & what) : std::runtime_error(what), code(c) {}
};

// hyper_parameter.hpp:35-43: same nine names, same defaults
struct HyperParameters : lfx_params
{
  HyperParameters() { lfx_default_params(this); }
  static HyperParameters LaunchYaml() { HyperParameters p; lfx_launch_yaml_params(&p); return p; }
};

// One field of sensor_msgs/PointCloud2::fields, as much as the callback reads of it
struct PointField { std::string name; uint32_t offset; uint8_t datatype; };

// Field lookup by name, as pcl::fromROSMsg does for PointXYZIR (lib/include/lidar_feature_library/point_type.hpp:83-86).
// RingIsAvailable (ring.cpp:36-44) becomes view.has_ring; is_dense is taken from the message.
// data_bytes: size of the message's data array (msg->data.size()); when given, a message whose width * height *
// point_step exceeds it is refused here instead of becoming an out-of-bounds read of the copy engine.
inline lfx_cloud_view MakeView(const void * data, uint32_t n_points, uint32_t point_step, const std::vector<PointField> & fields,
                               bool is_dense, bool device_memory = false, size_t data_bytes = 0)
{
  if (data_bytes != 0 && static_cast<uint64_t>(n_points) * point_step > data_bytes) {
    throw Error(LFX_E_BAD_LAYOUT, "width * height * point_step exceeds the size of the data array");
  }
  lfx_cloud_view v;
  std::memset(&v, 0, sizeof(v));
  v.data = data;
  v.n_points = n_points;
  v.point_step = point_step;
  v.is_dense = is_dense ? 1 : 0;
  v.memory = device_memory ? LFX_MEM_DEVICE : LFX_MEM_HOST;
  v.ring_datatype = LFX_RING_U16;
  int have = 0;
  for (const auto & f : fields) {
    if (f.name == "x") { v.off_x = f.offset; have |= 1; }
    else if (f.name == "y") { v.off_y = f.offset; have |= 2; }
    else if (f.name == "z") { v.off_z = f.offset; have |= 4; }
    else if (f.name == "ring") { v.off_ring = f.offset; v.ring_datatype = f.datatype; v.has_ring = 1; }
  }
  if (have != 7) { throw Error(LFX_E_BAD_LAYOUT, "PointCloud2 lacks x, y or z"); }
  return v;
}

class FeatureExtraction
{
public:
  explicit FeatureExtraction(const HyperParameters & params = HyperParameters(), int device = 0)
  {
    lfx_options opt;
    std::memset(&opt, 0, sizeof(opt));
    opt.device = device;
    const int rc = lfx_create(&params, &opt, &h_);
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(nullptr)); }
  }
  ~FeatureExtraction() { lfx_destroy(h_); }
  FeatureExtraction(const FeatureExtraction &) = delete;
  FeatureExtraction & operator=(const FeatureExtraction &) = delete;

  // One scan, synchronous: what the callback needs to fill scan_edge / scan_surface (16-byte x,y,z,1 points,
  // pinned host memory owned by the handle, valid until the next call) and colored_scan (labels).
  lfx_scan_output Extract(const lfx_cloud_view & scan)
  {
    lfx_scan_output out;
    Check(lfx_extract_scan(h_, &scan, &out));
    return out;
  }

  // Many independent scans, asynchronous (offline map building): results stay on the device.
  lfx_batch_result ExtractBatch(const std::vector<lfx_cloud_view> & scans)
  {
    lfx_batch_result res;
    Check(lfx_extract_batch(h_, scans.data(), static_cast<int>(scans.size()), &res));
    return res;
  }
  void Synchronize() { Check(lfx_synchronize(h_)); }
  // Ring ids of scan 0 of the last call that contributed nothing because the reference's per-ring code throws for
  // them (feature_extraction.cpp:154-156 logs a warning per such ring): lfx_ring_info.status == LFX_RING_SKIPPED.
  std::vector<int> SkippedRings(int max_rings = 128)
  {
    std::vector<lfx_ring_info> rings(static_cast<size_t>(max_rings));
    Check(lfx_fetch_rings(h_, rings.data()));
    std::vector<int> ids;
    for (int r = 0; r < max_rings; r++) { if (rings[r].status == LFX_RING_SKIPPED) { ids.push_back(r); } }
    return ids;
  }
  lfx_handle * handle() { return h_; }
  // cudaStream_t all work of this handle is enqueued on; chain collectives / timing events on it (lfx_stream).
  void * stream() const { return lfx_stream(h_); }

  // colored_scan of scan `scan` of the last batch (feature_extraction.cpp:153,161): 32-byte pcl::PointXYZRGB records,
  // i.e. PointCloud2.data of the message; needs lfx_options.want_sorted_src (use the two-argument constructor).
  FeatureExtraction(const HyperParameters & params, const lfx_options & opt)
  {
    const int rc = lfx_create(&params, &opt, &h_);
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(nullptr)); }
  }
  std::vector<uint8_t> ColoredScan(int scan = 0)
  {
    lfx_colored_result res;
    Check(lfx_color_batch(h_, &res));
    if (scan < 0 || scan >= res.n_scans) { throw Error(LFX_E_BAD_PARAM, "scan index out of range"); }
    std::vector<uint8_t> data(static_cast<size_t>(res.counts[scan]) * 32);
    Check(lfx_fetch_colored(h_, scan, data.data(), data.size()));
    return data;
  }

private:
  void Check(int rc) { if (rc != LFX_OK) { throw Error(rc, lfx_last_error(h_)); } }
  lfx_handle * h_ = nullptr;
};

// Offline batches from HOST memory, pipelined: two handles (two streams, two sets of device buffers) take the batches
// in turn, so that the device-to-host copy of batch k-1's features runs while batch k's host-to-device copy is on the
// other stream (PCIe is full duplex) and batch k's kernels hide under the upload of batch k+1. Submit(k) enqueues and
// returns; Collect() hands out the OLDEST batch in flight. Scans are independent (feature_extraction.cpp:92,173-175), so
// the results equal those of one handle called batch by batch. Input buffers of a batch must stay valid until its Collect.
class Pipeline
{
public:
  struct Output
  {
    std::vector<uint32_t> counts;    // [n_scans][2]
    std::vector<uint32_t> offsets;   // [n_scans + 1][2]
    uint32_t n_edge = 0, n_surface = 0;
  };
  explicit Pipeline(const HyperParameters & params = HyperParameters(), int device = 0) : fe0_(params, device), fe1_(params, device) {}
  void Submit(const std::vector<lfx_cloud_view> & scans)
  {
    if (in_flight_ == 2) { throw Error(LFX_E_STATE, "two batches in flight: Collect() first"); }
    n_scans_[head_] = scans.size();
    handle(head_).ExtractBatch(scans);
    head_ ^= 1;
    in_flight_++;
  }
  // features into caller memory (pinned for full PCIe speed, lfx_host_alloc_on): capacities in points
  Output Collect(float * edge_xyz, size_t edge_capacity, float * surface_xyz, size_t surface_capacity)
  {
    if (in_flight_ == 0) { throw Error(LFX_E_STATE, "no batch in flight"); }
    const int slot = in_flight_ == 2 ? head_ : head_ ^ 1;
    FeatureExtraction & fe = handle(slot);
    Output out;
    out.counts.resize(2 * n_scans_[slot]);
    out.offsets.resize(2 * (n_scans_[slot] + 1));
    int rc = lfx_fetch_counts(fe.handle(), out.counts.data(), out.offsets.data());
    if (rc == LFX_OK) { rc = lfx_batch_status(fe.handle()); }
    if (rc == LFX_OK) { rc = lfx_fetch_features(fe.handle(), edge_xyz, edge_capacity, surface_xyz, surface_capacity); }
    in_flight_--;
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(fe.handle())); }
    out.n_edge = out.offsets[2 * n_scans_[slot]];
    out.n_surface = out.offsets[2 * n_scans_[slot] + 1];
    return out;
  }
  int InFlight() const { return in_flight_; }
  FeatureExtraction & handle(int slot) { return (slot & 1) ? fe1_ : fe0_; }

private:
  FeatureExtraction fe0_, fe1_;
  size_t n_scans_[2] = {0, 0};
  int head_ = 0, in_flight_ = 0;
};

// fields + point_step of scan_edge / scan_surface / colored_scan (what pcl::toROSMsg derives, ros_msg.hpp:53-71)
struct TopicLayout
{
  std::vector<PointField> fields;
  uint32_t point_step = 0;
  explicit TopicLayout(int topic)
  {
    lfx_point_field f[4];
    uint32_t n = 0;
    if (lfx_topic_layout(topic, f, &n, &point_step) != LFX_OK) { throw Error(LFX_E_BAD_PARAM, "unknown topic"); }
    for (uint32_t k = 0; k < n; k++) { fields.push_back({f[k].name, f[k].offset, f[k].datatype}); }
  }
};

// The upstream converter (convert.py:171-212) on the device. Shares the extraction handle so that converted clouds
// feed ExtractBatch without leaving the GPU: Convert(raw clouds) -> View(i) -> FeatureExtraction::ExtractBatch.
class PointTypeConverter
{
public:
  explicit PointTypeConverter(FeatureExtraction & fe) : h_(fe.handle()) {}

  // Returns the per-cloud outcome; clouds for which the reference's callback raises have status != LFX_CONVERT_OK.
  lfx_convert_result Convert(const std::vector<lfx_raw_cloud> & clouds)
  {
    lfx_convert_result res;
    const int rc = lfx_convert_batch(h_, clouds.data(), static_cast<int>(clouds.size()), &res);
    if (rc != LFX_OK && rc != LFX_E_CONVERT) { throw Error(rc, lfx_last_error(h_)); }
    return res;
  }
  lfx_cloud_view View(int cloud)
  {
    lfx_cloud_view v;
    const int rc = lfx_converted_view(h_, cloud, &v);
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(h_)); }
    return v;
  }
  // views of every successfully converted cloud of the last Convert, in cloud order: the argument of ExtractBatch
  std::vector<lfx_cloud_view> Views(int n_clouds)
  {
    std::vector<lfx_cloud_view> v(static_cast<size_t>(n_clouds > 0 ? n_clouds : 1));
    int n = 0;
    const int rc = lfx_converted_views(h_, v.data(), n_clouds, &n);
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(h_)); }
    v.resize(static_cast<size_t>(n));
    return v;
  }
  // PointCloud2.data of /points_converted (convert.py:198-212: point_step 32, height 1, width = size / 32, is_dense)
  std::vector<uint8_t> Fetch(int cloud, uint32_t kept)
  {
    std::vector<uint8_t> data(static_cast<size_t>(kept) * 32);
    const int rc = lfx_fetch_converted(h_, cloud, data.data(), data.size());
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(h_)); }
    return data;
  }

private:
  lfx_handle * h_;
};

// The mapping package's accumulate step (mapping/include/lidar_feature_mapping/map.hpp:62-145) for the batched
// offline sequence: AddBatch(poses) does for every frame of the last extracted batch what MapBuilder::Callback does
// for one (scan_edge, pose) pair. The map stays on the device; Points() copies it out (x,y,z,1.0f per point).
class MapBuilder
{
public:
  explicit MapBuilder(FeatureExtraction & fe) : h_(fe.handle()) {}
  std::vector<uint8_t> AddBatch(const std::vector<lfx_pose> & poses)   // 1 where the frame was added
  {
    std::vector<uint8_t> selected(poses.size());
    uint64_t n = 0;
    Check(lfx_map_add_batch(h_, poses.data(), static_cast<int>(poses.size()), selected.data(), &n));
    return selected;
  }
  bool IsEmpty() { return Size() == 0; }
  uint64_t Size() { uint64_t n = 0; Check(lfx_map_size(h_, &n)); return n; }
  std::vector<float> Points()
  {
    std::vector<float> xyz(static_cast<size_t>(Size()) * 4);
    Check(lfx_map_fetch(h_, 0, xyz.size() / 4, xyz.data()));
    return xyz;
  }
  // sharded driver: the gate state after the frames before this rank's shard (lfx_map_gate over the gathered sizes)
  void SetState(bool map_empty, const lfx_pose * prev) { Check(lfx_map_set_state(h_, map_empty ? 1 : 0, prev)); }

private:
  void Check(int rc) { if (rc != LFX_OK) { throw Error(rc, lfx_last_error(h_)); } }
  lfx_handle * h_;
};

// The localization consumer's residual build (LOAMOptimizationProblem::Make, loam_optimization_problem.hpp:62-84, for
// maps held on the device): Edge / Surface return the Jacobian blocks and residuals of all features of a scan at once.
class LoamProblem
{
public:
  // maps: 16-byte x,y,z,1 points (pcl::PointXYZ), host or device memory
  LoamProblem(FeatureExtraction & fe, const float * edge_map, uint64_t n_edge, const float * surface_map, uint64_t n_surface,
              int n_neighbors = 15, bool device_memory = false)
  : h_(fe.handle()), k_(n_neighbors)
  {
    const int mem = device_memory ? LFX_MEM_DEVICE : LFX_MEM_HOST;
    Check(lfx_loc_set_map(h_, LFX_LOC_EDGE, edge_map, n_edge, mem));
    Check(lfx_loc_set_map(h_, LFX_LOC_SURFACE, surface_map, n_surface, mem));
  }
  // jacobians [n][3][7] (columns q_w,q_x,q_y,q_z,t_x,t_y,t_z), residuals [n][3]
  void Edge(const float * scan, uint32_t n, const lfx_pose & point_to_map, std::vector<double> & jacobians, std::vector<double> & residuals,
            bool device_memory = false)
  {
    jacobians.resize(static_cast<size_t>(n) * 21);
    residuals.resize(static_cast<size_t>(n) * 3);
    Check(lfx_loc_edge(h_, scan, n, device_memory ? LFX_MEM_DEVICE : LFX_MEM_HOST, &point_to_map, k_, jacobians.data(), residuals.data(), nullptr));
  }
  // jacobians [n][7], residuals [n]; `scan` is the surface scan AFTER its voxel down-sampling (surface.hpp:106-112)
  void Surface(const float * scan, uint32_t n, const lfx_pose & point_to_map, std::vector<double> & jacobians, std::vector<double> & residuals,
               bool device_memory = false)
  {
    jacobians.resize(static_cast<size_t>(n) * 7);
    residuals.resize(n);
    Check(lfx_loc_surface(h_, scan, n, device_memory ? LFX_MEM_DEVICE : LFX_MEM_HOST, &point_to_map, k_, jacobians.data(), residuals.data(), nullptr));
  }

private:
  void Check(int rc) { if (rc != LFX_OK) { throw Error(rc, lfx_last_error(h_)); } }
  lfx_handle * h_;
  int k_;
};

// One rank of the multi-GPU driver ("multi-GPU driver owns 8 handles + NCCL comm"; 8(e)): frames are
// sharded by index, rank g of G owning [g F / G, (g + 1) F / G). Per batch: fe.ExtractBatch(this rank's scans), then
// Exchange(); Fetch() (or Finish() for the device-side tables) whenever the frame-ordered global tables are needed.
class Shard
{
public:
  // one process (or thread) per rank: `unique_id` = the bytes rank 0 got from UniqueId()
  Shard(FeatureExtraction & fe, const std::vector<uint8_t> & unique_id, int rank, int world, uint64_t n_frames)
  {
    const int rc = lfx_shard_create(fe.handle(), unique_id.empty() ? nullptr : unique_id.data(), rank, world, n_frames, &s_);
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(fe.handle())); }
  }
  static std::vector<uint8_t> UniqueId()
  {
    std::vector<uint8_t> id(LFX_SHARD_ID_BYTES);
    const int rc = lfx_shard_unique_id(id.data());
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(nullptr)); }
    return id;
  }
  // all ranks in this process: one Shard per FeatureExtraction (distinct devices), rank = position
  static std::vector<Shard> Local(const std::vector<FeatureExtraction *> & fes, uint64_t n_frames)
  {
    std::vector<lfx_handle *> hs;
    for (FeatureExtraction * fe : fes) { hs.push_back(fe->handle()); }
    std::vector<lfx_shard *> out(fes.size(), nullptr);
    const int rc = lfx_shard_create_local(hs.data(), static_cast<int>(hs.size()), n_frames, out.data());
    if (rc != LFX_OK) { throw Error(rc, lfx_last_error(hs.empty() ? nullptr : hs[0])); }
    std::vector<Shard> shards;
    for (lfx_shard * s : out) { shards.push_back(Shard(s)); }
    return shards;
  }
  Shard(Shard && o) noexcept : s_(o.s_) { o.s_ = nullptr; }
  Shard(const Shard &) = delete;
  Shard & operator=(const Shard &) = delete;
  ~Shard() { lfx_shard_destroy(s_); }

  static void Range(uint64_t n_frames, int rank, int world, uint64_t & first, uint64_t & last)
  {
    if (lfx_shard_range(n_frames, rank, world, &first, &last) != LFX_OK) { throw Error(LFX_E_BAD_PARAM, "rank outside the world"); }
  }
  void Exchange() { Check(lfx_shard_exchange(s_)); }
  lfx_shard_result Finish() { lfx_shard_result r; Check(lfx_shard_finish(s_, &r)); return r; }
  // counts [n_frames][2] and offsets [n_frames + 1][2] of the last exchange, in frame order (synchronises)
  void Fetch(std::vector<uint32_t> & counts, std::vector<uint64_t> & offsets, uint64_t n_frames)
  {
    counts.resize(2 * n_frames);
    offsets.resize(2 * (n_frames + 1));
    Check(lfx_shard_fetch(s_, counts.data(), offsets.data()));
  }
  bool UsesPeerStores() const { int p = 0; lfx_shard_info(s_, &p, nullptr); return p != 0; }

private:
  explicit Shard(lfx_shard * s) : s_(s) {}
  void Check(int rc) { if (rc != LFX_OK) { throw Error(rc, lfx_shard_last_error(s_)); } }
  lfx_shard * s_ = nullptr;
};

}  // namespace lfx
#endif  // LFX_HPP_
