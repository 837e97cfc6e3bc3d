/*
 * lfx.h — C ABI of the B200-native LiDAR feature extraction path.
 *
 * Drop-in boundary for ONE path of tier4/lidar_feature_extraction: the body of
 * FeatureExtraction::Callback (extraction/app/feature_extraction.cpp:92-171), i.e. lines 110-157
 * (ring extraction, curvature, labelling, masks, edge/surface gathering) plus the PointXYZ
 * conversion of lines 163-164. Everything is `extern "C"`, plain pointers and sizes; no C++ or
 * torch types cross this boundary. INTEGRATION.md shows the rclcpp-side stub that calls it.
 *
 * All "replaces" citations are file:line under the reference repository.
 *
 * The implementation is CUDA-only (sm_100a). There is no CPU fallback: without a usable device
 * lfx_create() fails with LFX_E_CUDA.
 */
#ifndef LFX_H_
#define LFX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFX_VERSION_MAJOR 0
#define LFX_VERSION_MINOR 1

/* ------------------------------------------------------------------ status codes
 * The reference signals these conditions by RCLCPP_ERROR + rclcpp::shutdown()
 * (feature_extraction.cpp:96-108) or asserts (hyper_parameter.hpp:45-53). */
enum {
  LFX_OK = 0,
  LFX_E_BAD_PARAM = 1,  /* hyper_parameter.hpp:45-53 asserts; or outside the supported envelope */
  LFX_E_NOT_DENSE = 2,  /* feature_extraction.cpp:96-101 */
  LFX_E_NO_RING = 3,    /* feature_extraction.cpp:103-108, RingIsAvailable ring.cpp:36-44 */
  LFX_E_BAD_LAYOUT = 4, /* PointCloud2 view that pcl::fromROSMsg could not map either */
  LFX_E_CAPACITY = 5,   /* ring id >= max_rings (see lfx_options), or an output buffer that is too small */
  LFX_E_CUDA = 6,
  LFX_E_STATE = 7,      /* call order violated (e.g. fetch before extract) */
  LFX_E_CONVERT = 8     /* lfx_convert_batch: at least one cloud failed where the reference converter raises;
                           per-cloud reasons in lfx_convert_result.status */
};

/* ------------------------------------------------------------------ labels
 * PointLabel, extraction/include/lidar_feature_extraction/point_label.hpp:32-42. */
enum {
  LFX_LABEL_DEFAULT = 0,
  LFX_LABEL_EDGE = 1,
  LFX_LABEL_EDGE_NEIGHBOR = 2,
  LFX_LABEL_SURFACE = 3,
  LFX_LABEL_SURFACE_NEIGHBOR = 4,
  LFX_LABEL_OUT_OF_RANGE = 5,
  LFX_LABEL_OCCLUDED = 6,
  LFX_LABEL_PARALLEL_BEAM = 7,
  LFX_LABEL_NONE = 255 /* point of a ring that contributes nothing (sparse or skipped, see below) */
};

/* per-ring status in lfx_ring_info */
enum {
  LFX_RING_OK = 0,
  LFX_RING_SPARSE = 1,  /* fewer than padding+1 points: RemoveSparseRings, ring.cpp:46-59 */
  LFX_RING_SKIPPED = 2, /* the reference would throw std::invalid_argument and WARN,
                           feature_extraction.cpp:154-156 (too short for the convolution / sectors,
                           or two adjacent points with zero XY norm, math.cpp:40-42) */
  LFX_RING_TOO_LONG = 3 /* transient: a ring above the on-chip capacity before k_extract_rings_big has extracted it */
};

/* ------------------------------------------------------------------ parameters
 * Replaces HyperParameters (hyper_parameter.hpp:32-65): same nine fields, same meaning.
 * ROS parameter names: convolution_padding, neighbor_degree_threshold, distance_diff_threshold,
 * parallel_beam_min_range_ratio, edge_threshold, surface_threshold, min_range, max_range, n_blocks. */
typedef struct lfx_params {
  int padding;
  double neighbor_degree_threshold;
  double distance_diff_threshold;
  double parallel_beam_min_range_ratio;
  double edge_threshold;
  double surface_threshold;
  double min_range;
  double max_range;
  int n_blocks;
} lfx_params;

/* compiled defaults, hyper_parameter.hpp:35-43 */
void lfx_default_params(lfx_params *out);
/* deployed set, lidar_feature_launch/config/lidar_feature_extraction.param.yaml:3-10 */
void lfx_launch_yaml_params(lfx_params *out);

/* Sizing and diagnostics knobs that have no counterpart in the reference. Zero = default. */
typedef struct lfx_options {
  int device;            /* CUDA device ordinal */
  int max_ring_points;   /* longest ring held on chip; default 4096, at most 8192 (larger values are clamped, and
                          * lowered to what the device's shared memory holds). Longer
                          * rings are not refused: they run on the unbounded per-ring kernel (lfx_big.cuh) */
  int max_rings;         /* ring ids must be < max_rings; default 128, max 4096 */
  int want_sorted_src;   /* also produce the ring-sorted -> source index map (4 B/point) */
  int want_curvature;    /* also produce per-point curvature, f64 (8 B/point; test/diagnostic mode) */
  int force_order_path;  /* testing: 0 auto, 1 force key sort, 2 force exact comparator sort */
  void *stream;          /* cudaStream_t to enqueue on; NULL = a stream owned by the handle */
  int use_graph;         /* 1 (default when 0 is passed with graph_default) capture batches in CUDA graphs; -1 disables */
} lfx_options;

/* ------------------------------------------------------------------ input
 * Replaces GetPointCloud<PointXYZIR>(PointCloud2) (lib/include/lidar_feature_library/ros_msg.hpp:73-79,
 * point_type.hpp:62-86): fields are located by the caller (by name, as pcl::fromROSMsg does) and
 * passed as byte offsets. Deployed layout (point_type_converter/convert.py:134-145): point_step 32,
 * x@0 y@4 z@8 intensity@16 (f32), ring@20 (u16). Points need not be organised or ring-ordered. */
enum { LFX_MEM_HOST = 0, LFX_MEM_DEVICE = 1 };
enum { LFX_RING_U8 = 2, LFX_RING_U16 = 4, LFX_RING_U32 = 6 }; /* sensor_msgs/PointField datatype ids */

typedef struct lfx_cloud_view {
  const void *data;       /* PointCloud2.data */
  uint32_t n_points;      /* width * height */
  uint32_t point_step;
  uint32_t off_x, off_y, off_z, off_ring;
  uint8_t ring_datatype;  /* LFX_RING_U16 in the deployed layout */
  uint8_t has_ring;       /* 0 => LFX_E_NO_RING */
  uint8_t is_dense;       /* 0 => LFX_E_NOT_DENSE */
  uint8_t memory;         /* LFX_MEM_HOST (pageable or pinned) or LFX_MEM_DEVICE */
} lfx_cloud_view;

/* ------------------------------------------------------------------ results */
typedef struct lfx_ring_info {
  uint32_t count;       /* points carrying this ring id */
  uint32_t offset;      /* first position of the ring inside its scan's ring-sorted arrays */
  uint32_t n_edge;
  uint32_t n_surface;
  uint32_t status;      /* LFX_RING_* */
  uint32_t order_path;  /* diagnostics: 0 rotated-monotone, 1 key sort, 2 exact comparator sort */
} lfx_ring_info;

/* Device-side view of the last batch. Pointers stay valid until the next lfx_extract_batch /
 * lfx_destroy on the handle. All arrays are on the handle's device.
 *
 * Feature clouds replace what the reference publishes on scan_edge / scan_surface after
 * ToPointXYZ (feature_extraction.cpp:163-166): 16-byte points x,y,z,1.0f (pcl::PointXYZ layout).
 * Within a scan: rings ascending by id, inside a ring ascending sorted index
 * (GetIndicesByValue, lib/include/lidar_feature_library/algorithm.hpp:50-62). Scans concatenated
 * in batch order; scan s owns [offsets[2s], offsets[2s]+counts[2s]) of edge_xyz and
 * [offsets[2s+1], ...+counts[2s+1]) of surface_xyz.
 *
 * Where this differs from the reference AS SHIPPED (consumers of the topics, read this):
 *  - ORDER of the points inside scan_edge / scan_surface / colored_scan. The reference appends ring after ring in the
 *    iteration order of a std::unordered_map (feature_extraction.cpp:120), i.e. in an unspecified order; here rings
 *    are ascending by id. The per-ring content (the SET of points with each label) is the same
 *    (tests/test_oracle_vs_ref.py::test_feature_sets_match_the_verbatim_reference_independent_of_order);
 *    localization and mapping build kd-trees / concatenate these clouds and do not depend on the order.
 *  - TIES. Argsort (algorithm.hpp:65-71) and SortByAtan2 (ring.hpp:101-112) use std::sort, which leaves the order of
 *    equal keys to the library. Here equal curvatures are walked in index order and equal polar angles keep source
 *    order, i.e. the results are those of the reference built with a stable (value, index) sort
 *    (oracle/_ref/libref_stable.so: the bit-exact parity bar). Exact curvature ties do not occur on float32 sensor
 *    data; on constructed inputs that have them, the shipped reference's own labels depend on its libstdc++.
 *
 * Host buffers behind LFX_MEM_HOST views: lfx_extract_batch returns before their host-to-device copies have
 * finished. They must stay valid and unmodified until the NEXT-BUT-ONE lfx_extract_batch on the handle has returned
 * (two descriptor slots), or until lfx_synchronize / any lfx_fetch_* of this batch has returned. */
typedef struct lfx_batch_result {
  int n_scans;
  uint64_t total_points;
  const float *d_edge_xyz;        /* [sum n_edge][4] */
  const float *d_surface_xyz;     /* [sum n_surface][4] */
  const uint32_t *d_counts;       /* [n_scans][2]   (n_edge, n_surface) */
  const uint32_t *d_offsets;      /* [n_scans+1][2] exclusive prefix of d_counts (last row = totals) */
  const uint8_t *d_labels;        /* [total_points] final PointLabel per point in ring-sorted order */
  const uint32_t *d_sorted_src;   /* [total_points] or NULL: source index inside the scan */
  const double *d_curvature;      /* [total_points] or NULL */
  const lfx_ring_info *d_rings;   /* [n_scans][max_rings] */
  const uint64_t *d_point_base;   /* [n_scans+1] first position of each scan in the per-point arrays */
  int max_rings;
} lfx_batch_result;

typedef struct lfx_handle lfx_handle;

/* ------------------------------------------------------------------ lifecycle */
/* Replaces the FeatureExtraction node's constructor state (feature_extraction.cpp:65-88, 173-175):
 * parameters are immutable after creation. */
int lfx_create(const lfx_params *params, const lfx_options *options, lfx_handle **out);
void lfx_destroy(lfx_handle *h);
/* Human-readable description of the last non-OK status (h may be NULL for create failures). */
const char *lfx_last_error(const lfx_handle *h);
int lfx_get_params(const lfx_handle *h, lfx_params *out);
int lfx_device(const lfx_handle *h);

/* ------------------------------------------------------------------ extraction
 * Replaces feature_extraction.cpp:110-157 + 163-164 for `n_scans` independent PointCloud2 messages
 * (the callback is const and stateless, :92, so scans batch freely). Work is enqueued on the handle's
 * stream; the call returns once everything is enqueued (host views are copied H2D asynchronously, so
 * host buffers must stay valid until lfx_synchronize or a fetch). `out` may be NULL. */
int lfx_extract_batch(lfx_handle *h, const lfx_cloud_view *scans, int n_scans, lfx_batch_result *out);
int lfx_synchronize(lfx_handle *h);

/* Per-batch status after completion: returns LFX_OK or LFX_E_CAPACITY (first offending scan/ring in
 * lfx_last_error). Synchronises. */
int lfx_batch_status(lfx_handle *h);

/* D2H helpers (synchronise the handle's stream). Buffers are caller-owned.
 * counts: [n_scans][2] u32; offsets: [n_scans+1][2] u32. */
int lfx_fetch_counts(lfx_handle *h, uint32_t *counts, uint32_t *offsets);
/* edge/surface: capacity in points (16 B each); copies the whole batch's concatenated clouds. */
int lfx_fetch_features(lfx_handle *h, float *edge_xyz, size_t edge_capacity_points,
                       float *surface_xyz, size_t surface_capacity_points);
/* per-point arrays, total_points long each (NULL to skip). sorted_src / curvature need the
 * corresponding lfx_options flag. */
int lfx_fetch_points(lfx_handle *h, uint8_t *labels, uint32_t *sorted_src, double *curvature);
int lfx_fetch_rings(lfx_handle *h, lfx_ring_info *rings /* [n_scans][max_rings] */);

/* Convenience for a ROS-callback-shaped caller: one PointCloud2 in, two pcl::PointXYZ payloads out
 * (pinned host memory owned by the handle, valid until the next call). Synchronous. */
typedef struct lfx_scan_output {
  const float *edge_xyz;     /* [n_edge][4] */
  const float *surface_xyz;  /* [n_surface][4] */
  uint32_t n_edge, n_surface;
  const uint8_t *labels;     /* [n_points], ring-sorted order (source for colored_scan) */
  const uint32_t *sorted_src;/* [n_points] or NULL */
  uint32_t n_points;
} lfx_scan_output;
int lfx_extract_scan(lfx_handle *h, const lfx_cloud_view *scan, lfx_scan_output *out);

/* LabelToColor, extraction/src/color_points.cpp:39-68 (colored_scan is derived on the host from
 * the label bytes; it is a depth-1 debug topic, feature_extraction.cpp:77-78). rgb[3]. */
int lfx_label_to_color(uint8_t label, uint8_t *rgb);

/* ------------------------------------------------------------------ upstream converter
 * Replaces PointTypeConverter.callback (point_type_converter/point_type_converter/convert.py:183-212): a raw
 * driver PointCloud2 - any field set, any sensor_msgs/PointField datatype, either byte order - becomes the
 * deployed 32-byte layout (make_fields, convert.py:137-145: x,y,z,padding,intensity f32 @0..19, ring u16 @20,
 * bytes 22..31 zero, little-endian, is_dense) with the all-zero returns removed (nonzero, convert.py:165-166),
 * in the input's point order. The reference's struct-format behaviour is kept to the letter: a FLOAT32 field
 * 'padding' at offset 12 is appended, fields are read in offset order at their effective offsets
 * (create_point_format, convert.py:69-81), the zero test looks at the first three fields in offset order, and
 * the fields named x, y, z, padding, intensity, ring are packed positionally into '<fffffH'. */
typedef struct lfx_point_field {   /* sensor_msgs/PointField */
  const char *name;
  uint32_t offset;
  uint8_t datatype;                /* 1 INT8, 2 UINT8, 3 INT16, 4 UINT16, 5 INT32, 6 UINT32, 7 FLOAT32, 8 FLOAT64 */
  uint32_t count;                  /* ignored, as in the reference (one value per field) */
} lfx_point_field;

typedef struct lfx_raw_cloud {     /* the parts of sensor_msgs/PointCloud2 the converter reads */
  const void *data;
  uint64_t data_bytes;             /* must be a multiple of point_step (convert.py:91-92) */
  uint32_t point_step;
  const lfx_point_field *fields;
  uint32_t n_fields;
  uint8_t is_bigendian;
  uint8_t memory;                  /* LFX_MEM_HOST or LFX_MEM_DEVICE */
} lfx_raw_cloud;

/* per-cloud outcome: where the reference's callback raises, and why */
enum {
  LFX_CONVERT_OK = 0,
  LFX_CONVERT_E_SIZE = 1,          /* data size is not a multiple of point_step (convert.py:91-92) */
  LFX_CONVERT_E_DATATYPE = 2,      /* datatype id outside 1..8 (KeyError in create_point_format) */
  LFX_CONVERT_E_LAYOUT = 3,        /* fields run past point_step: struct.unpack size mismatch (convert.py:90-97) */
  LFX_CONVERT_E_FEW_FIELDS = 4,    /* fewer than three fields: nonzero() indexes c[0..2] (convert.py:165-166) */
  LFX_CONVERT_E_FIELD_COUNT = 5,   /* kept points do not carry exactly six retained values (struct.pack) */
  LFX_CONVERT_E_OVERFLOW = 6,      /* a finite value does not fit the 'f' format */
  LFX_CONVERT_E_RING_TYPE = 7,     /* the sixth retained field is a float: 'H' needs an integer */
  LFX_CONVERT_E_RING_RANGE = 8     /* 'H' format requires 0 <= ring <= 65535 */
};

/* Result of the last lfx_convert_batch. Arrays are host memory owned by the handle; d_points is device memory.
 * Cloud c owns kept[c] converted points at d_points + 32 * point_base[c] (point_base is the prefix sum of the
 * INPUT point counts, so every cloud has room for all of its points). Valid until the next lfx_convert_batch. */
typedef struct lfx_convert_result {
  int n_clouds;
  const uint8_t *d_points;
  const uint64_t *point_base;      /* [n_clouds + 1] */
  const uint32_t *kept;            /* [n_clouds] = width of the converted cloud (0 for failed clouds) */
  const uint32_t *status;          /* [n_clouds] LFX_CONVERT_* */
} lfx_convert_result;

/* Converts n_clouds independent clouds in one launch. Synchronous (the caller needs the widths). Returns LFX_OK,
 * or LFX_E_CONVERT when at least one cloud failed (the others are converted). */
int lfx_convert_batch(lfx_handle *h, const lfx_raw_cloud *clouds, int n_clouds, lfx_convert_result *out);
/* Device time of the last batch's conversion kernel (CUDA events on the handle's stream). */
int lfx_last_convert_ms(lfx_handle *h, float *ms);
/* View of converted cloud `cloud` (device memory, deployed layout) ready for lfx_extract_batch. */
int lfx_converted_view(lfx_handle *h, int cloud, lfx_cloud_view *out);
/* All successfully converted clouds of the last lfx_convert_batch at once, in cloud order (clouds for which the
 * reference's callback raises are left out, as nothing is published for them): out[capacity], *n_out views. This is
 * the array to hand to lfx_extract_batch for the deployed chain /points_raw -> converter -> extraction. The ring
 * ids of these clouds are also kept as a compact by-product of the conversion, which lfx_extract_batch uses when it
 * has to bucket them (the clouds must not be modified in between). */
int lfx_converted_views(lfx_handle *h, lfx_cloud_view *out, int capacity, int *n_out);
/* Copies converted cloud `cloud` (32 * kept bytes = PointCloud2.data of /points_converted) to host memory. */
int lfx_fetch_converted(lfx_handle *h, int cloud, void *dst, size_t capacity_bytes);

/* ------------------------------------------------------------------ colored_scan + message layouts
 * Replaces ColorPointsByLabel (extraction/include/lidar_feature_extraction/color_points.hpp:60-74) as the node
 * uses it (feature_extraction.cpp:153,161,168): all points of the rings that contribute to the outputs (not
 * sparse, not skipped), rings ascending, ring-sorted order, as 32-byte pcl::PointXYZRGB records
 * x,y,z,1.0f,{b,g,r,a=255},12 zero bytes. Needs lfx_options.want_sorted_src. Synchronous. Arrays of the result
 * are host memory owned by the handle; d_points is device memory: scan s owns counts[s] points at
 * d_points + 32 * point_base[s]. Valid until the next lfx_extract_batch / lfx_color_batch. */
typedef struct lfx_colored_result {
  int n_scans;
  const uint8_t *d_points;
  const uint64_t *point_base;   /* [n_scans + 1] */
  const uint32_t *counts;       /* [n_scans] */
} lfx_colored_result;
int lfx_color_batch(lfx_handle *h, lfx_colored_result *out);
int lfx_fetch_colored(lfx_handle *h, int scan, void *dst, size_t capacity_bytes);

/* PointCloud2 layout of the node's three output topics, i.e. what pcl::toROSMsg makes of pcl::PointXYZ
 * (scan_edge, scan_surface: x,y,z FLOAT32 @0,4,8, point_step 16) and pcl::PointXYZRGB (colored_scan: x,y,z
 * @0,4,8 and rgb FLOAT32 @16, point_step 32) - lib/include/lidar_feature_library/ros_msg.hpp:53-71. The
 * library's feature / colored buffers are exactly PointCloud2.data of these messages (height 1, width n,
 * little-endian, is_dense, frame "lidar_feature_base_link", stamp of the input: feature_extraction.cpp:159-166). */
enum { LFX_TOPIC_SCAN_EDGE = 0, LFX_TOPIC_SCAN_SURFACE = 1, LFX_TOPIC_COLORED_SCAN = 2 };
int lfx_topic_layout(int topic, lfx_point_field *fields /* capacity 4 */, uint32_t *n_fields, uint32_t *point_step);

/* ------------------------------------------------------------------ mapping accumulate
 * Replaces, for the batched offline sequence, MapBuilder::Callback + Map::TransformAdd of the mapping package
 * (mapping/include/lidar_feature_mapping/map.hpp:104-127, :68-74; thresholds :92-93): frame i of the last
 * extracted batch is added to the map iff its scan_edge cloud is not empty and (the map is empty or the pose
 * moved >= 1 m or rotated |dq.vec| >= 0.1 since the last ADDED frame: PoseDiffIsSufficientlySmall, :50-60); an
 * added cloud is transformed by its pose (GetIsometry3d, lib/src/ros_msg.cpp:33-38; pcl::transformPointCloud in
 * double) and appended. The gate is sequential and runs on the host; the transform + append run on the device.
 * The map lives on the device as 16-byte x,y,z,1.0f points in frame order and persists across batches. */
typedef struct lfx_pose {      /* geometry_msgs/Pose */
  double position[3];          /* x, y, z */
  double orientation[4];       /* x, y, z, w */
} lfx_pose;
/* poses: one per scan of the last lfx_extract_batch. selected_out (optional, [n_scans]): 1 where the frame was
 * added. Synchronous. */
int lfx_map_add_batch(lfx_handle *h, const lfx_pose *poses, int n_poses, uint8_t *selected_out, uint64_t *map_points_out);
int lfx_map_size(lfx_handle *h, uint64_t *n_points_out);
int lfx_map_fetch(lfx_handle *h, uint64_t first, uint64_t n_points, float *xyz /* [n_points][4] */);
int lfx_map_clear(lfx_handle *h);
/* The gate alone, on the host (no device work): MapBuilder::Callback's decisions (map.hpp:117-131) for a sequence of
 * frames given their poses and scan_edge sizes. *map_empty_io / *prev_io carry the builder's state in and out
 * (prev_io is only read when *map_empty_io == 0). selected: [n] 0/1. A rank of the sharded driver runs it over the
 * frames before its shard (sizes from the count all-gather) and installs the result with lfx_map_set_state, so that
 * its part of the map equals the corresponding part of the sequential map. */
int lfx_map_gate(const lfx_pose *poses, const uint32_t *n_edge, int n, int *map_empty_io, lfx_pose *prev_io, uint8_t *selected);
int lfx_map_set_state(lfx_handle *h, int map_empty, const lfx_pose *prev);
/* PoseDiffIsSufficientlySmall (map.hpp:50-60) on two poses: 1 small, 0 not (host arithmetic, exposed for tests). */
int lfx_pose_diff_is_small(const lfx_pose *pose0, const lfx_pose *pose1, double translation_threshold, double rotation_threshold);

/* ------------------------------------------------------------------ memory helpers */
/* Pinned host memory so that H2D/D2H run at full PCIe speed. */
void *lfx_host_alloc(size_t bytes);
void lfx_host_free(void *p);
/* Pinned host memory on the NUMA node of the handle's GPU (sysfs numa_node of its PCI device, else the memory affinity NVML reports; set_mempolicy around the
 * allocation). *numa_node_out (may be NULL): the node the pages were bound to, -1 if the platform gave none. Use it
 * for the scan buffers of a rank that feeds its GPU from host memory; free with lfx_host_free. */
void *lfx_host_alloc_on(lfx_handle *h, size_t bytes, int *numa_node_out);
int lfx_device_alloc(lfx_handle *h, size_t bytes, void **out);
int lfx_device_free(lfx_handle *h, void *p);
int lfx_memcpy_h2d(lfx_handle *h, void *dst_device, const void *src_host, size_t bytes);
int lfx_memcpy_d2h(lfx_handle *h, void *dst_host, const void *src_device, size_t bytes);

/* The cudaStream_t every kernel and copy of this handle is enqueued on (the one given in lfx_options, or the
 * stream the handle created). A caller that chains its own work (the sharded driver's NCCL all-gather of the
 * per-scan counts, timing events) must enqueue it on this stream or order against it. */
void *lfx_stream(const lfx_handle *h);

/* ------------------------------------------------------------------ instrumentation */
/* Kernels launched / graph launches issued by this handle so far (bench.py's gpu_launches). */
uint64_t lfx_kernel_launch_count(const lfx_handle *h);
/* Device time of the last batch's stages, measured with CUDA events on the handle's stream:
 * ms[0]=layout probe, ms[1]=sector kernel on regular scans, ms[2]=ring bucketing of the other scans (list +
 * hist + plan + scatter + ring probe), ms[3]=sector kernel on bucketed rings, ms[4]=per-ring (sort) kernel,
 * ms[5]=pack. Only when timing was enabled. enabled = 1: the batch is launched kernel by kernel with event records
 * in between (no graph); enabled = 2: the batch runs as its CUDA graph, as in production, with the events as
 * event-record nodes of that graph (what bench.py reports); 0: off. */
#define LFX_N_STAGES 6
int lfx_set_stage_timing(lfx_handle *h, int enabled);
int lfx_last_stage_ms(lfx_handle *h, float *ms /* [LFX_N_STAGES] */);

/* Which path the scans of the last batch took (synchronises). A scan is "regular" when its points are
 * in sensor firing order with a fixed ring period and every ring is a rotated monotone sequence of polar
 * angles; regular scans run on the sector kernel straight from the PointCloud2 payload. Every other scan
 * (returns dropped by the converter, arbitrary point order, or a failed check) is bucketed by ring id first;
 * its rings that are rotated monotone sequences run on the same sector kernel through the bucket's index
 * list, the remaining rings are sorted and processed by the per-ring kernel. Results are identical. */
typedef struct lfx_batch_stats {
  uint32_t fast_rings[3];   /* rings of regular scans handed to the sector kernel, per positions-per-lane class (10, 11, 12) */
  uint32_t general_scans;   /* scans that were bucketed by ring id */
  uint32_t general_rings;   /* rings processed by the per-ring (sort) kernel */
  uint32_t indexed_rings[3];/* bucketed rings handed to the sector kernel, per class */
} lfx_batch_stats;
int lfx_last_batch_stats(lfx_handle *h, lfx_batch_stats *out);

/* ------------------------------------------------------------------ synthetic scans
 * Deterministic generator for the BASELINE.json sensor shapes (test/bench input, not part of the
 * reference). Emits the deployed 32-byte layout in sensor firing order (azimuth-major, ring-minor,
 * clockwise from a per-scan random start azimuth). */
enum { LFX_WORLD_ROOM = 0, LFX_WORLD_TUNNEL = 1 };
typedef struct lfx_synth_spec {
  int n_rings, n_cols;
  float elev_lo_deg, elev_hi_deg; /* ring 0 .. ring n_rings-1 */
  int world;                      /* LFX_WORLD_* */
  float range_noise;              /* metres, gaussian sigma */
  float dropout_prob;             /* probability that a return is missing (points removed) */
  float dropout_burst;            /* mean burst length in columns (>= 1) */
  float near_prob;                /* probability of a spurious < min_range return */
  uint64_t seed;
} lfx_synth_spec;
/* named shapes: "vlp16" 16x1800, "hdl32" 32x2170, "hdl64" 64x2048 (tunnel), "os128" 128x2048 */
int lfx_synth_named(const char *name, lfx_synth_spec *out);
/* host generator: out capacity n_rings*n_cols*32 bytes; writes actual point count */
int lfx_synth_scan_host(const lfx_synth_spec *spec, uint64_t frame, void *out, uint32_t *n_points_out);
/* device generator (no drop-outs): n_scans consecutive frames, each n_rings*n_cols*32 bytes, contiguous */
int lfx_synth_batch_device(lfx_handle *h, const lfx_synth_spec *spec, uint64_t first_frame, int n_scans,
                           void *d_out);

/* ------------------------------------------------------------------ localization residual build
 * First slice of the consumer's hot loop: what Edge<..>::Make (localization/include/lidar_feature_localization/
 * edge.hpp:88-124) and Surface<..>::MakeFromDownsampled (surface.hpp:116-139) compute for every feature of a scan in
 * every iteration of the optimiser: the n_neighbors nearest points of the edge / surface map (kdtree.cpp:42-55), the
 * line (mean + principal axis of the neighbours) or plane (least squares X w = -1) through them, and the feature's
 * Jacobian block and residual. The neighbour search is exact and returns the same index lists as the reference's
 * kd-tree (ascending distance; nanoflann's L2 over doubles). The voxel down-sampling of the surface scan
 * (surface.hpp:106-112) is PCL's and stays with the caller: pass the down-sampled scan.
 * Layouts: map and scan points are 16-byte x,y,z,1.0f (pcl::PointXYZ, what lfx_extract_* emit and lfx_map_* hold);
 * point_to_map: rotation as quaternion (x,y,z,w) + translation; edge Jacobians [n][3][7] row major with columns
 * (q_w, q_x, q_y, q_z, t_x, t_y, t_z) as MakeEdgeJacobianRow (edge.cpp:64-73), residuals [n][3]; surface Jacobians
 * [n][7] (MakeJacobianRow, surface.hpp:84-92), residuals [n]; neighbors (optional) [n][n_neighbors] map indices.
 * Synchronous; outputs are host arrays. */
#define LFX_LOC_EDGE 0
#define LFX_LOC_SURFACE 1
int lfx_loc_set_map(lfx_handle *h, int kind, const float *xyz4, uint64_t n_points, int memory /* LFX_MEM_* */);
int lfx_loc_edge(lfx_handle *h, const float *scan_xyz4, uint32_t n, int memory, const lfx_pose *point_to_map, int n_neighbors,
                 double *jacobians, double *residuals, uint32_t *neighbors);
int lfx_loc_surface(lfx_handle *h, const float *scan_xyz4, uint32_t n, int memory, const lfx_pose *point_to_map, int n_neighbors,
                    double *jacobians, double *residuals, uint32_t *neighbors);
int lfx_loc_release(lfx_handle *h);   /* frees the maps and work buffers (also done by lfx_destroy) */

/* ------------------------------------------------------------------ multi-GPU driver (SURVEY.md 8(b) "Threading", 8(e))
 * Scans are independent (the reference's callback is const and stateless, feature_extraction.cpp:92,173-175): a
 * sequence of n_frames scans is sharded by frame index, rank g of G owning frames [g F / G, (g + 1) F / G) on its own
 * GPU and handle. No point or feature crosses GPUs; the one exchange is that every rank's per-scan counts
 * (n_edge, n_surface) reach every rank, which derives where each frame's clouds start in the frame-ordered
 * concatenation. One lfx_shard per rank, either one process per GPU (lfx_shard_create; NCCL rendezvous through a
 * unique id) or all ranks in one process (lfx_shard_create_local). NCCL (loaded at run time) sets the group up; the
 * per-batch exchange itself is a one-CTA kernel that stores the counts into every peer's receive buffer through
 * peer-mapped memory over NVLink and raises a flag, so no collective kernel has to wait for SMs behind the persistent
 * extraction kernel (set LFX_SHARD_EXCHANGE=nccl to use ncclAllGather instead; it is also the fallback when the
 * buffers cannot be peer-mapped).
 *
 * Per batch:  lfx_extract_batch(h, this rank's scans) ; lfx_shard_exchange(s) ;  ... next batch ...
 * and lfx_shard_finish / lfx_shard_fetch whenever the global tables are needed (the consumer side of an exchange is
 * enqueued lazily: at the next exchange, or here). Everything is enqueued on the handle's stream. */
#define LFX_SHARD_ID_BYTES 128
typedef struct lfx_shard lfx_shard;
typedef struct lfx_shard_result {
  uint64_t n_frames;
  uint64_t first_frame, last_frame;  /* this rank's shard [first, last) */
  const uint32_t *d_counts_all;      /* [n_frames][2] (n_edge, n_surface) in frame order, on this rank's device */
  const uint64_t *d_offsets_all;     /* [n_frames + 1][2] exclusive prefix (last row = totals) */
} lfx_shard_result;
int lfx_shard_range(uint64_t n_frames, int rank, int world, uint64_t *first, uint64_t *last);
/* rank 0: LFX_SHARD_ID_BYTES bytes to hand to every rank by whatever means the host program has (ncclGetUniqueId) */
int lfx_shard_unique_id(void *id_out);
/* one process (or thread) per rank; blocks until all `world` ranks have called it (ncclCommInitRank) */
int lfx_shard_create(lfx_handle *h, const void *unique_id, int rank, int world, uint64_t n_frames, lfx_shard **out);
/* all ranks in the calling process: handles[world] on distinct devices -> out[world] (ncclCommInitAll) */
int lfx_shard_create_local(lfx_handle **handles, int world, uint64_t n_frames, lfx_shard **out);
/* publish the counts of the handle's last batch (which must be this rank's shard) to every rank; asynchronous */
int lfx_shard_exchange(lfx_shard *s);
/* global tables of the last exchange, valid once the handle's stream has reached this point (lfx_synchronize) */
int lfx_shard_finish(lfx_shard *s, lfx_shard_result *out);
/* the same, copied to the host (either pointer may be NULL); synchronises; LFX_E_STATE if a peer never published */
int lfx_shard_fetch(lfx_shard *s, uint32_t *counts_all /* [n_frames][2] */, uint64_t *offsets_all /* [n_frames+1][2] */);
int lfx_shard_info(const lfx_shard *s, int *uses_peer_stores, int *nccl_ranks);
const char *lfx_shard_last_error(const lfx_shard *s);
void lfx_shard_destroy(lfx_shard *s);

#ifdef __cplusplus
}
#endif
#endif /* LFX_H_ */
