// Reference-backed oracle: links the reference's own extraction sources (compiled in place
// from /root/reference, see oracle/Makefile) behind a tiny C entry point.
//
// TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// reference arm may load this library; the product path never does.
//
// ref_extract_scan() is the per-scan body of FeatureExtraction::Callback
// (extraction/app/feature_extraction.cpp:110-157) with the ROS/PCL I/O removed: that file itself
// cannot be compiled here (rclcpp, pcl_conversions). Every call below is to the reference's
// unmodified function of the same name. Rings are emitted in ascending ring id (the reference
// iterates an unordered_map, feature_extraction.cpp:120, so its order is unspecified).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "lidar_feature_extraction/color_points.hpp"
#include "lidar_feature_extraction/curvature.hpp"
#include "lidar_feature_extraction/index_range.hpp"
#include "lidar_feature_extraction/label.hpp"
#include "lidar_feature_extraction/math.hpp"
#include "lidar_feature_extraction/neighbor.hpp"
#include "lidar_feature_extraction/occlusion.hpp"
#include "lidar_feature_extraction/out_of_range.hpp"
#include "lidar_feature_extraction/parallel_beam.hpp"
#include "lidar_feature_extraction/point_label.hpp"
#include "lidar_feature_extraction/range.hpp"
#include "lidar_feature_extraction/ring.hpp"

#include "lidar_feature_library/algorithm.hpp"
#include "lidar_feature_library/degree_to_radian.hpp"
#include "lidar_feature_library/point_type.hpp"

extern "C" {

struct ref_params
{
  int padding;
  double neighbor_degree_threshold;
  double distance_diff_threshold;
  double parallel_beam_min_range_ratio;
  double edge_threshold;
  double surface_threshold;
  double min_range;
  double max_range;
  int n_blocks;
};

// points: n x 32-byte PointXYZIR records (x,y,z,1.0f,intensity,ring u16).
// Outputs (caller-allocated, capacity n each unless noted):
//   ring_ids[n_rings_cap], ring_sizes[n_rings_cap], ring_skipped[n_rings_cap]: per surviving ring
//       (after RemoveSparseRings), ascending ring id; skipped = the ring threw invalid_argument.
//   sorted_src[n]: source point index of each point in (ring asc, angle-sorted) order; rings that
//       were removed as sparse are absent; skipped rings ARE present (their labels are 255).
//   labels[n], curvature[n]: per sorted point (curvature 0/labels 255 for skipped rings).
//   edge_idx / surface_idx: positions into sorted_src order of Edge / Surface points.
// Returns the number of sorted points written, or -1 on capacity error.
int ref_extract_scan(
  const void * points, int n, const ref_params * prm,
  int n_rings_cap, int * n_rings_out, int * ring_ids, int * ring_sizes, int * ring_skipped,
  int * sorted_src, std::uint8_t * labels, double * curvature,
  int * n_edge_out, int * edge_idx, int * n_surface_out, int * surface_idx)
{
  pcl::PointCloud<PointXYZIR>::Ptr input_cloud(new pcl::PointCloud<PointXYZIR>());
  input_cloud->points.resize(n);
  if (n > 0) {
    std::memcpy(static_cast<void *>(input_cloud->points.data()), points, sizeof(PointXYZIR) * n);
  }

  const EdgeLabel edge_label(prm->padding, prm->edge_threshold);
  const SurfaceLabel surface_label(prm->padding, prm->surface_threshold);

  // feature_extraction.cpp:114-118
  auto rings_unordered = ExtractAngleSortedRings(*input_cloud);
  RemoveSparseRings(rings_unordered, prm->padding + 1);
  const std::map<int, std::vector<int>> rings(rings_unordered.begin(), rings_unordered.end());

  int n_rings = 0, n_sorted = 0, n_edge = 0, n_surface = 0;
  for (const auto & [ring, indices] : rings) {
    if (n_rings >= n_rings_cap) {return -1;}
    const int base = n_sorted;
    const int size = static_cast<int>(indices.size());
    for (int i = 0; i < size; i++) {
      sorted_src[base + i] = indices[i];
      labels[base + i] = 255;
      curvature[base + i] = 0.;
    }
    n_sorted += size;
    ring_ids[n_rings] = ring;
    ring_sizes[n_rings] = size;
    ring_skipped[n_rings] = 0;

    // feature_extraction.cpp:121-156
    try {
      const MappedPoints<PointXYZIR> ref_points(input_cloud, indices);
      const double radian_threshold = DegreeToRadian(prm->neighbor_degree_threshold);
      const NeighborCheckXY<PointXYZIR> is_neighbor(ref_points, radian_threshold);
      const Range<PointXYZIR> range(ref_points);

      std::vector<PointLabel> lab = InitLabels(ref_points.size());
      const std::vector<double> ranges = range(0, range.size());
      const std::vector<double> curv = CalcCurvature(ranges, prm->padding);
      const PaddedIndexRange index_range(range.size(), prm->n_blocks, prm->padding);

      AssignLabel(lab, curv, is_neighbor, index_range, edge_label, surface_label);

      LabelOccludedPoints(lab, is_neighbor, range, prm->padding, prm->distance_diff_threshold);
      LabelOutOfRange(lab, range, prm->min_range, prm->max_range);
      LabelParallelBeamPoints(lab, range, prm->parallel_beam_min_range_ratio);

      const std::vector<size_t> e = GetIndicesByValue(lab, PointLabel::Edge);
      const std::vector<size_t> s = GetIndicesByValue(lab, PointLabel::Surface);
      for (int i = 0; i < size; i++) {
        labels[base + i] = static_cast<std::uint8_t>(lab[i]);
        curvature[base + i] = curv[i];
      }
      for (size_t i : e) {edge_idx[n_edge++] = base + static_cast<int>(i);}
      for (size_t i : s) {surface_idx[n_surface++] = base + static_cast<int>(i);}
    } catch (const std::invalid_argument &) {
      ring_skipped[n_rings] = 1;
    }
    n_rings++;
  }
  *n_rings_out = n_rings;
  *n_edge_out = n_edge;
  *n_surface_out = n_surface;
  return n_sorted;
}

// colored_scan of one scan: the same loop, with the reference's own ColorPointsByLabel / LabelToColor
// (color_points.hpp:60-74, color_points.cpp:39-68) appended per ring exactly where feature_extraction.cpp:153 does it
// (inside the try block: a ring that throws contributes nothing). Rings ascending like ref_extract_scan.
// xyz[3 * n], rgb[3 * n] (r, g, b as the reference's MakeXYZRGB assigns them). Returns the number of points.
int ref_color_scan(const void * points, int n, const ref_params * prm, float * xyz, std::uint8_t * rgb)
{
  pcl::PointCloud<PointXYZIR>::Ptr input_cloud(new pcl::PointCloud<PointXYZIR>());
  input_cloud->points.resize(n);
  if (n > 0) {
    std::memcpy(static_cast<void *>(input_cloud->points.data()), points, sizeof(PointXYZIR) * n);
  }
  const EdgeLabel edge_label(prm->padding, prm->edge_threshold);
  const SurfaceLabel surface_label(prm->padding, prm->surface_threshold);
  auto rings_unordered = ExtractAngleSortedRings(*input_cloud);
  RemoveSparseRings(rings_unordered, prm->padding + 1);
  const std::map<int, std::vector<int>> rings(rings_unordered.begin(), rings_unordered.end());
  pcl::PointCloud<pcl::PointXYZRGB>::Ptr colored_cloud(new pcl::PointCloud<pcl::PointXYZRGB>());
  for (const auto & [ring, indices] : rings) {
    (void)ring;
    try {
      const MappedPoints<PointXYZIR> ref_points(input_cloud, indices);
      const double radian_threshold = DegreeToRadian(prm->neighbor_degree_threshold);
      const NeighborCheckXY<PointXYZIR> is_neighbor(ref_points, radian_threshold);
      const Range<PointXYZIR> range(ref_points);
      std::vector<PointLabel> lab = InitLabels(ref_points.size());
      const std::vector<double> ranges = range(0, range.size());
      const std::vector<double> curv = CalcCurvature(ranges, prm->padding);
      const PaddedIndexRange index_range(range.size(), prm->n_blocks, prm->padding);
      AssignLabel(lab, curv, is_neighbor, index_range, edge_label, surface_label);
      LabelOccludedPoints(lab, is_neighbor, range, prm->padding, prm->distance_diff_threshold);
      LabelOutOfRange(lab, range, prm->min_range, prm->max_range);
      LabelParallelBeamPoints(lab, range, prm->parallel_beam_min_range_ratio);
      *colored_cloud += *ColorPointsByLabel<PointXYZIR>(ref_points, lab);
    } catch (const std::invalid_argument &) {
    }
  }
  int m = 0;
  for (const auto & p : colored_cloud->points) {
    xyz[3 * m] = p.x; xyz[3 * m + 1] = p.y; xyz[3 * m + 2] = p.z;
    rgb[3 * m] = p.r; rgb[3 * m + 1] = p.g; rgb[3 * m + 2] = p.b;
    m++;
  }
  return m;
}

// Piecewise entry points used to pin the C restatement function by function.
int ref_polar_less(float ax, float ay, float bx, float by)
{
  struct P {float x, y;};
  return AHasSmallerPolarAngleThanB<P>{}(P{ax, ay}, P{bx, by}) ? 1 : 0;
}

int ref_curvature(const double * range, int n, int padding, double * out)
{
  try {
    const std::vector<double> r(range, range + n);
    const std::vector<double> c = CalcCurvature(r, padding);
    std::copy(c.begin(), c.end(), out);
    return 0;
  } catch (const std::invalid_argument &) {
    return 1;
  }
}

int ref_boundaries(int size, int n_blocks, int padding, int * out)
{
  try {
    const PaddedIndexRange r(size, n_blocks, padding);
    for (int j = 0; j < n_blocks; j++) {out[j] = r.Begin(j);}
    out[n_blocks] = r.End(n_blocks - 1);
    return 0;
  } catch (const std::invalid_argument &) {
    return 1;
  }
}

// returns 0/1, or -1 if CalcRadian throws (both norms zero)
int ref_is_neighbor(float x1, float y1, float x2, float y2, double radian_threshold)
{
  try {
    return CalcRadian(x1, y1, x2, y2) < radian_threshold ? 1 : 0;
  } catch (const std::invalid_argument &) {
    return -1;
  }
}

const char * ref_variant()
{
#ifdef LFX_REF_STABLE
  return "stable";
#else
  return "verbatim";
#endif
}

}  // extern "C"
