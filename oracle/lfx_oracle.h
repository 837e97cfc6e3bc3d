/*
 * lfx_oracle.h — CPU restatement of tier4/lidar_feature_extraction's per-scan extraction path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` leg may load liblfx_oracle.so. The CUDA product path never
 * calls into it and has no CPU fallback.
 *
 * Parity status: PINNED. Every function here is checked (tests/test_oracle_golden.py) against the
 * known-answer vectors of the reference's own gtest files, and (tests/test_oracle_vs_ref.py)
 * bit-for-bit against the reference's sources compiled in place (oracle/_ref/libref_*.so).
 *
 * All citations are file:line under /root/reference.
 */
#ifndef LFX_ORACLE_H_
#define LFX_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PointLabel, extraction/include/lidar_feature_extraction/point_label.hpp:32-42 */
enum {
  LFXO_DEFAULT = 0,
  LFXO_EDGE = 1,
  LFXO_EDGE_NEIGHBOR = 2,
  LFXO_SURFACE = 3,
  LFXO_SURFACE_NEIGHBOR = 4,
  LFXO_OUT_OF_RANGE = 5,
  LFXO_OCCLUDED = 6,
  LFXO_PARALLEL_BEAM = 7,
  LFXO_SKIPPED = 255 /* ring threw std::invalid_argument: contributes nothing (feature_extraction.cpp:154-156) */
};

/* HyperParameters, extraction/include/lidar_feature_extraction/hyper_parameter.hpp:32-65 */
typedef struct lfxo_params {
  int padding;
  double neighbor_degree_threshold;
  double distance_diff_threshold;
  double parallel_beam_min_range_ratio;
  double edge_threshold;
  double surface_threshold;
  double min_range;
  double max_range;
  int n_blocks;
} lfxo_params;

/* PointCloud2 payload view (fields looked up by name upstream; lib ros_msg.hpp:73-79). */
typedef struct lfxo_cloud {
  const void *data;
  int n_points;
  int point_step;
  int off_x, off_y, off_z, off_ring;
  int ring_datatype; /* sensor_msgs PointField: 2=UINT8, 4=UINT16, 6=UINT32 */
} lfxo_cloud;

/* ---- piecewise restatements (each is pinned by the reference's gtest vectors) ---- */

/* AHasSmallerPolarAngleThanB<PointXYZIR>, ring.hpp:54-99 (float arithmetic, as for PCL points). */
int lfxo_polar_less_f32(float ax, float ay, float bx, float by);
/* same comparator instantiated on double members (the reference's tests use a double Point). */
int lfxo_polar_less_f64(double ax, double ay, double bx, double by);
/* SortByAtan2, ring.hpp:101-112: sorts idx[0..n) so the polar angle of (x[idx],y[idx]) ascends. */
void lfxo_sort_by_polar_angle_f32(const float *x, const float *y, int *idx, int n);
void lfxo_sort_by_polar_angle_f64(const double *x, const double *y, int *idx, int n);
/* XYNorm math.hpp:36-39 */
double lfxo_xy_norm(double x, double y);
/* CalcRadian math.cpp:34-46; returns 0 and writes *out, or 1 if both norms are zero (throws there). */
int lfxo_calc_radian(double x1, double y1, double x2, double y2, double *out);
/* IsNeighborXY neighbor.hpp:44-48; returns 0/1, or -1 where the reference throws. */
int lfxo_is_neighbor(float x1, float y1, float x2, float y2, double radian_threshold);
/* DegreeToRadian lib/include/lidar_feature_library/degree_to_radian.hpp:34-37 */
double lfxo_degree_to_radian(double degree);
/* MakeWeight curvature.cpp:36-42: out[2*padding+1] */
void lfxo_make_weight(int padding, double *out);
/* Convolution1D convolution.cpp:35-66; returns 1 where the reference throws (n < n_weight). */
int lfxo_convolution_1d(const double *input, int n, const double *weight, int n_weight, double *out);
/* CalcCurvature curvature.cpp:44-50; returns 1 where the reference throws. */
int lfxo_curvature(const double *range, int n, int padding, double *out);
/* IndexRange::Boundary index_range.cpp:60-66 on [start,end) with n_blocks; out[n_blocks+1].
 * returns 1 where the IndexRange ctor throws (index_range.cpp:35-40). */
int lfxo_index_range(int start, int end, int n_blocks, int *out);
/* PaddedIndexRange index_range.hpp:59-66 */
int lfxo_padded_index_range(int size, int n_blocks, int padding, int *out);
/* Argsort algorithm.hpp:65-71 with the (value, index) tie-break. */
void lfxo_argsort(const double *values, int n, int *out);
/* FillFromLeft / FillFromRight / FillNeighbors fill.hpp:40-117.
 * link[i] (0 <= i < n-1) = is_neighbor(i, i+1). Return 1 where the reference throws. */
int lfxo_fill_from_left(uint8_t *labels, const uint8_t *link, int n, int begin, int end, uint8_t label);
int lfxo_fill_from_right(uint8_t *labels, const uint8_t *link, int n, int begin, int end, uint8_t label);
int lfxo_fill_neighbors(uint8_t *labels, const uint8_t *link, int n, int index, int padding, uint8_t label);
/* EdgeLabel::Assign label.hpp:72-95, SurfaceLabel::Assign label.hpp:113-134 on one sector view. */
void lfxo_edge_assign(uint8_t *labels, const double *curvature, const uint8_t *link, int n, int padding, double threshold);
void lfxo_surface_assign(uint8_t *labels, const double *curvature, const uint8_t *link, int n, int padding, double threshold);
/* FromLeft / FromRight / LabelOccludedPoints occlusion.hpp:37-91 */
void lfxo_occlusion_from_left(uint8_t *labels, const uint8_t *link, const double *range, int n, int padding, double d);
void lfxo_occlusion_from_right(uint8_t *labels, const uint8_t *link, const double *range, int n, int padding, double d);
/* LabelOutOfRange out_of_range.hpp:36-48 */
void lfxo_out_of_range(uint8_t *labels, const double *range, int n, double min_range, double max_range);
/* LabelParallelBeamPoints parallel_beam.hpp:36-51 */
void lfxo_parallel_beam(uint8_t *labels, const double *range, int n, double ratio);
/* LabelToColor color_points.cpp:39-68: rgb[3]; returns 1 for an invalid label. */
int lfxo_label_to_color(uint8_t label, uint8_t *rgb);

/* ---- one ring, feature_extraction.cpp:121-151 on angle-sorted points ----
 * x,y: the ring's points in sorted order. labels[n], curvature[n] written.
 * Returns 0, or 1 if the ring would throw std::invalid_argument (then labels are LFXO_SKIPPED). */
int lfxo_extract_ring(const float *x, const float *y, int n, const lfxo_params *prm,
                      uint8_t *labels, double *curvature);

/* ---- one scan, feature_extraction.cpp:110-157 ----
 * Same output convention as oracle/ref_driver.cpp:ref_extract_scan (rings ascending by id). */
int lfxo_extract_scan(const lfxo_cloud *cloud, const lfxo_params *prm,
                      int n_rings_cap, int *n_rings_out, int *ring_ids, int *ring_sizes, int *ring_skipped,
                      int *sorted_src, uint8_t *labels, double *curvature,
                      int *n_edge_out, int *edge_idx, int *n_surface_out, int *surface_idx);

/* ---- a batch of equally laid-out scans over `n_threads` pthreads, frames round-robin (the only
 * parallelism the reference admits: scans are independent, feature_extraction.cpp:92). Only counts
 * are returned: this is the CPU timing leg. counts[2*s] = n_edge, counts[2*s+1] = n_surface. */
int lfxo_extract_batch_counts(const lfxo_cloud *clouds, int n_scans, const lfxo_params *prm,
                              int n_threads, int *counts);

#ifdef __cplusplus
}
#endif
#endif /* LFX_ORACLE_H_ */
