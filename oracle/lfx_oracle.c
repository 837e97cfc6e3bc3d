/*
 * lfx_oracle.c — plain-C restatement of the reference's per-scan extraction path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see lfx_oracle.h). Parity status: PINNED against the
 * reference's own gtest vectors and against its sources compiled in place (oracle/_ref).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (x86-64 baseline: no FMA, like the reference's
 * build, extraction/CMakeLists.txt:6-7). Citations are file:line under /root/reference.
 */
#include "lfx_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ ring ingest */

/* ring.hpp:54-99, evaluated in T = float (PCL points) or double (the reference's test Point).
 * `lena`, `lenb` and `det` are T-typed expressions widened to double afterwards. */
#define LFXO_POLAR_LESS(T, NAME)                                        \
  int NAME(T ax, T ay, T bx, T by)                                      \
  {                                                                     \
    if (ax == bx && ay == by) { return 0; }         /* ring.hpp:62-64 */ \
    const double lena = (T)(ax * ax + ay * ay);     /* ring.hpp:66 */    \
    const double lenb = (T)(bx * bx + by * by);     /* ring.hpp:67 */    \
    if (lena == 0) {                                /* ring.hpp:69-76 */ \
      if (by == 0) { return bx < 0; }                                   \
      return by > 0;                                                    \
    }                                                                   \
    if (lenb == 0) { return ay < 0; }               /* ring.hpp:78-80 */ \
    if (ay == 0) { return (ax >= 0) && (by >= 0); } /* ring.hpp:82-84 */ \
    if (by == 0) { return !((bx >= 0) && (ay >= 0)); } /* :86-88 */      \
    if ((T)(ay * by) > 0) {                         /* ring.hpp:91-94 */ \
      const double det = (T)((T)(ax * by) - (T)(ay * bx));              \
      return det > 0;                                                   \
    }                                                                   \
    return ay < 0;                                  /* ring.hpp:96 */    \
  }

LFXO_POLAR_LESS(float, lfxo_polar_less_f32)
LFXO_POLAR_LESS(double, lfxo_polar_less_f64)

/* SortByAtan2 ring.hpp:101-112. The reference uses std::sort; any comparison sort gives the same
 * permutation when the comparator is a strict weak order without equivalent elements. A bottom-up
 * merge sort is used so that equivalent elements keep source order (deterministic). */
#define LFXO_SORT_POLAR(T, NAME, LESS)                                              \
  void NAME(const T *x, const T *y, int *idx, int n)                                \
  {                                                                                 \
    if (n < 2) { return; }                                                          \
    int *tmp = (int *)malloc(sizeof(int) * (size_t)n);                              \
    int *src = idx, *dst = tmp;                                                     \
    for (int w = 1; w < n; w *= 2) {                                                \
      for (int lo = 0; lo < n; lo += 2 * w) {                                       \
        int mid = lo + w < n ? lo + w : n;                                          \
        int hi = lo + 2 * w < n ? lo + 2 * w : n;                                   \
        int a = lo, b = mid, k = lo;                                                \
        while (a < mid && b < hi) {                                                 \
          /* take from the right run only if it is strictly smaller */              \
          if (LESS(x[src[b]], y[src[b]], x[src[a]], y[src[a]])) { dst[k++] = src[b++]; } \
          else { dst[k++] = src[a++]; }                                             \
        }                                                                           \
        while (a < mid) { dst[k++] = src[a++]; }                                    \
        while (b < hi) { dst[k++] = src[b++]; }                                     \
      }                                                                             \
      int *t = src; src = dst; dst = t;                                             \
    }                                                                               \
    if (src != idx) { memcpy(idx, src, sizeof(int) * (size_t)n); }                  \
    free(tmp);                                                                      \
  }

LFXO_SORT_POLAR(float, lfxo_sort_by_polar_angle_f32, lfxo_polar_less_f32)
LFXO_SORT_POLAR(double, lfxo_sort_by_polar_angle_f64, lfxo_polar_less_f64)

/* ------------------------------------------------------------------ math */

double lfxo_xy_norm(double x, double y) { return sqrt(x * x + y * y); } /* math.hpp:36-39 */

double lfxo_degree_to_radian(double degree) { return degree * M_PI / 180.0; } /* degree_to_radian.hpp:34-37 */

int lfxo_calc_radian(double x1, double y1, double x2, double y2, double *out) /* math.cpp:34-46 */
{
  const double dot = x1 * x2 + y1 * y2;
  const double norm1 = lfxo_xy_norm(x1, y1);
  const double norm2 = lfxo_xy_norm(x2, y2);
  if (norm1 == 0 && norm2 == 0) { return 1; } /* throws std::invalid_argument, math.cpp:40-42 */
  *out = acos(dot / (norm1 * norm2));
  return 0;
}

int lfxo_is_neighbor(float x1, float y1, float x2, float y2, double radian_threshold) /* neighbor.hpp:44-48 */
{
  double rad;
  if (lfxo_calc_radian(x1, y1, x2, y2, &rad)) { return -1; }
  return rad < radian_threshold; /* NaN compares false */
}

/* ------------------------------------------------------------------ curvature */

void lfxo_make_weight(int padding, double *out) /* curvature.cpp:36-42 */
{
  for (int i = 0; i < 2 * padding + 1; i++) { out[i] = 1.; }
  out[padding] = -2. * padding;
}

int lfxo_convolution_1d(const double *input, int n, const double *weight, int n_weight, double *out)
{
  if (n < n_weight) { return 1; } /* convolution.cpp:39-43 */
  const int padding = (n_weight - 1) / 2;
  const int conv = n - padding * 2;
  for (int i = 0; i < padding; i++) { out[i] = 0.; }
  for (int i = 0; i < conv; i++) {
    double sum = 0.; /* InnerProduct math.hpp:43-53: left to right from 0.0 */
    for (int k = 0; k < n_weight; k++) { sum += input[i + k] * weight[k]; }
    out[padding + i] = sum;
  }
  for (int i = 0; i < padding; i++) { out[conv + padding + i] = 0.; }
  return 0;
}

int lfxo_curvature(const double *range, int n, int padding, double *out) /* curvature.cpp:44-50 */
{
  double weight[2 * 64 + 1];
  if (padding < 1 || padding > 64) { return 1; }
  lfxo_make_weight(padding, weight);
  if (lfxo_convolution_1d(range, n, weight, 2 * padding + 1, out)) { return 1; }
  for (int i = 0; i < n; i++) { out[i] = out[i] * out[i]; }
  return 0;
}

/* ------------------------------------------------------------------ sectors */

int lfxo_index_range(int start, int end, int n_blocks, int *out)
{
  if (end - start < n_blocks) { return 1; } /* index_range.cpp:35-40 */
  const double s = (double)start, e = (double)end, n = (double)n_blocks;
  for (int j = 0; j <= n_blocks; j++) {
    out[j] = (int)(s * (1. - j / n) + e * j / n); /* index_range.cpp:62-65 */
  }
  return 0;
}

int lfxo_padded_index_range(int size, int n_blocks, int padding, int *out) /* index_range.hpp:59-66 */
{
  return lfxo_index_range(padding, size - padding, n_blocks, out);
}

/* ------------------------------------------------------------------ argsort */

typedef struct { double v; int i; } lfxo_vi;

static int lfxo_vi_cmp(const void *pa, const void *pb)
{
  const lfxo_vi *a = (const lfxo_vi *)pa, *b = (const lfxo_vi *)pb;
  if (a->v < b->v) { return -1; }
  if (b->v < a->v) { return 1; }
  return (a->i > b->i) - (a->i < b->i); /* index tie-break (north_star; test_algorithm.cpp:44-48) */
}

void lfxo_argsort(const double *values, int n, int *out) /* algorithm.hpp:65-71 */
{
  lfxo_vi *t = (lfxo_vi *)malloc(sizeof(lfxo_vi) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) { t[i].v = values[i]; t[i].i = i; }
  qsort(t, (size_t)n, sizeof(lfxo_vi), lfxo_vi_cmp);
  for (int i = 0; i < n; i++) { out[i] = t[i].i; }
  free(t);
}

/* ------------------------------------------------------------------ fill */

int lfxo_fill_from_left(uint8_t *labels, const uint8_t *link, int n, int begin, int end, uint8_t label)
{
  if (end > n) { return 1; }   /* fill.hpp:50-53 */
  if (begin < 0) { return 1; } /* fill.hpp:55-58 */
  for (int i = begin; i < end - 1; i++) { /* fill.hpp:60-66 */
    labels[i] = label;
    if (!link[i]) { return 0; }
  }
  labels[end - 1] = label; /* fill.hpp:67 */
  return 0;
}

int lfxo_fill_from_right(uint8_t *labels, const uint8_t *link, int n, int begin, int end, uint8_t label)
{
  if (end >= n) { return 1; }   /* fill.hpp:80-84 */
  if (begin < -1) { return 1; } /* fill.hpp:86-89 */
  for (int i = end; i > begin + 1; i--) { /* fill.hpp:91-97 */
    labels[i] = label;
    if (!link[i - 1]) { return 0; }
  }
  labels[begin + 1] = label; /* fill.hpp:98 */
  return 0;
}

int lfxo_fill_neighbors(uint8_t *labels, const uint8_t *link, int n, int index, int padding, uint8_t label)
{
  const int lo = index - padding - 1 > -1 ? index - padding - 1 : -1; /* fill.hpp:112 */
  const int hi = index + 1 + padding < n ? index + 1 + padding : n;   /* fill.hpp:113 */
  if (lfxo_fill_from_right(labels, link, n, lo, index, label)) { return 1; }
  return lfxo_fill_from_left(labels, link, n, index, hi, label);
}

/* ------------------------------------------------------------------ selection */

void lfxo_edge_assign(uint8_t *labels, const double *curvature, const uint8_t *link, int n, int padding,
                      double threshold) /* label.hpp:72-95 */
{
  int *order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  lfxo_argsort(curvature, n, order);
  for (int k = n - 1; k >= 0; k--) { /* reverse walk, label.hpp:87 */
    const int index = order[k];
    if (!(labels[index] == LFXO_DEFAULT && curvature[index] >= threshold)) { continue; }
    lfxo_fill_neighbors(labels, link, n, index, padding, LFXO_EDGE_NEIGHBOR);
    labels[index] = LFXO_EDGE;
  }
  free(order);
}

void lfxo_surface_assign(uint8_t *labels, const double *curvature, const uint8_t *link, int n, int padding,
                         double threshold) /* label.hpp:113-134 */
{
  int *order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  lfxo_argsort(curvature, n, order);
  for (int k = 0; k < n; k++) {
    const int index = order[k];
    if (!(labels[index] == LFXO_DEFAULT && curvature[index] <= threshold)) { continue; }
    lfxo_fill_neighbors(labels, link, n, index, padding, LFXO_SURFACE_NEIGHBOR);
    labels[index] = LFXO_SURFACE;
  }
  free(order);
}

/* ------------------------------------------------------------------ masks */

void lfxo_occlusion_from_left(uint8_t *labels, const uint8_t *link, const double *range, int n, int padding,
                              double d) /* occlusion.hpp:37-57 */
{
  for (int i = 0; i < n - padding - 1; i++) {
    if (!link[i]) { continue; }
    if (range[i + 1] > range[i] + d) {
      lfxo_fill_from_left(labels, link, n, i + 1, i + padding + 2, LFXO_OCCLUDED);
    }
  }
}

void lfxo_occlusion_from_right(uint8_t *labels, const uint8_t *link, const double *range, int n, int padding,
                               double d) /* occlusion.hpp:59-79 */
{
  for (int i = n - 1; i >= padding + 1; i--) {
    if (!link[i - 1]) { continue; }
    if (range[i - 1] > range[i] + d) {
      lfxo_fill_from_right(labels, link, n, i - padding - 2, i - 1, LFXO_OCCLUDED);
    }
  }
}

void lfxo_out_of_range(uint8_t *labels, const double *range, int n, double min_range, double max_range)
{ /* out_of_range.hpp:36-48 with IsInInclusiveRange range.hpp:40-43 */
  for (int i = 0; i < n; i++) {
    if (!(min_range <= range[i] && range[i] <= max_range)) { labels[i] = LFXO_OUT_OF_RANGE; }
  }
}

void lfxo_parallel_beam(uint8_t *labels, const double *range, int n, double ratio) /* parallel_beam.hpp:36-51 */
{
  for (int i = 1; i < n - 1; i++) {
    const float ratio1 = (float)(fabs(range[i - 1] - range[i]) / range[i]); /* narrowed to float, :44 */
    const float ratio2 = (float)(fabs(range[i + 1] - range[i]) / range[i]); /* :45 */
    if ((double)ratio1 > ratio && (double)ratio2 > ratio) { labels[i] = LFXO_PARALLEL_BEAM; }
  }
}

int lfxo_label_to_color(uint8_t label, uint8_t *rgb) /* color_points.cpp:39-68 */
{
  static const uint8_t table[8][3] = {
    {255, 255, 255}, {255, 0, 0}, {255, 63, 0}, {255, 0, 0},
    {255, 63, 0}, {127, 127, 127}, {255, 0, 255}, {0, 255, 0}};
  if (label > 7) { return 1; }
  memcpy(rgb, table[label], 3);
  return 0;
}

/* ------------------------------------------------------------------ one ring */

int lfxo_extract_ring(const float *x, const float *y, int n, const lfxo_params *prm,
                      uint8_t *labels, double *curvature)
{
  const int P = prm->padding, B = prm->n_blocks;
  int rc = 1;
  for (int i = 0; i < n; i++) { labels[i] = LFXO_SKIPPED; curvature[i] = 0.; }

  double *range = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  uint8_t *link = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
  uint8_t *lab = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1); /* InitLabels label.hpp:56-59 */
  int *bnd = (int *)malloc(sizeof(int) * (size_t)(B + 1));

  if (n < 2) { goto done; } /* NeighborCheckXY ctor neighbor.hpp:71-75 */
  for (int i = 0; i < n; i++) { range[i] = lfxo_xy_norm(x[i], y[i]); } /* range.hpp:52-65 */
  if (lfxo_curvature(range, n, P, curvature)) { goto done; }   /* convolution.cpp:39-43 */
  if (lfxo_padded_index_range(n, B, P, bnd)) { goto done; }    /* index_range.cpp:35-40 */
  for (int j = 0; j < B; j++) {
    if (bnd[j + 1] - bnd[j] < 2) { goto done; } /* Slice -> NeighborCheckXY ctor, label.hpp:159 */
  }
  {
    /* Every adjacent pair is evaluated by the occlusion passes (occlusion.hpp:45,67) once
     * n >= 2P+1, so one pair with both XY norms zero makes CalcRadian throw (math.cpp:40-42). */
    const double theta = lfxo_degree_to_radian(prm->neighbor_degree_threshold);
    for (int i = 0; i < n - 1; i++) {
      const int r = lfxo_is_neighbor(x[i], y[i], x[i + 1], y[i + 1], theta);
      if (r < 0) { goto done; }
      link[i] = (uint8_t)r;
    }
  }
  for (int j = 0; j < B; j++) { /* AssignLabel label.hpp:153-163: views clipped to the sector */
    const int b = bnd[j], m = bnd[j + 1] - bnd[j];
    lfxo_edge_assign(lab + b, curvature + b, link + b, m, P, prm->edge_threshold);
    lfxo_surface_assign(lab + b, curvature + b, link + b, m, P, prm->surface_threshold);
  }
  /* feature_extraction.cpp:135-138: masks overwrite afterwards, whole ring */
  lfxo_occlusion_from_left(lab, link, range, n, P, prm->distance_diff_threshold);
  lfxo_occlusion_from_right(lab, link, range, n, P, prm->distance_diff_threshold);
  lfxo_out_of_range(lab, range, n, prm->min_range, prm->max_range);
  lfxo_parallel_beam(lab, range, n, prm->parallel_beam_min_range_ratio);
  memcpy(labels, lab, (size_t)n);
  rc = 0;
done:
  if (rc) { for (int i = 0; i < n; i++) { curvature[i] = 0.; } }
  free(range); free(link); free(lab); free(bnd);
  return rc;
}

/* ------------------------------------------------------------------ one scan */

static inline float lfxo_load_f32(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }

static inline int lfxo_load_ring(const uint8_t *p, int datatype)
{
  if (datatype == 2) { return p[0]; }
  if (datatype == 6) { uint32_t v; memcpy(&v, p, 4); return (int)v; }
  uint16_t v; memcpy(&v, p, 2); return v; /* UINT16: point_type.hpp:66,85 */
}

typedef struct { int ring; int idx; } lfxo_ri;

static int lfxo_ri_cmp(const void *pa, const void *pb)
{
  const lfxo_ri *a = (const lfxo_ri *)pa, *b = (const lfxo_ri *)pb;
  if (a->ring != b->ring) { return (a->ring > b->ring) - (a->ring < b->ring); }
  return (a->idx > b->idx) - (a->idx < b->idx);
}

int lfxo_extract_scan(const lfxo_cloud *cloud, const lfxo_params *prm,
                      int n_rings_cap, int *n_rings_out, int *ring_ids, int *ring_sizes, int *ring_skipped,
                      int *sorted_src, uint8_t *labels, double *curvature,
                      int *n_edge_out, int *edge_idx, int *n_surface_out, int *surface_idx)
{
  const int n = cloud->n_points;
  const uint8_t *data = (const uint8_t *)cloud->data;
  int n_rings = 0, n_sorted = 0, n_edge = 0, n_surface = 0, rc = 0;

  /* MakePointIndices ring.hpp:114-125: bucket by ring keeping source order */
  lfxo_ri *ri = (lfxo_ri *)malloc(sizeof(lfxo_ri) * (size_t)(n > 0 ? n : 1));
  float *x = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float *y = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float *rx = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float *ry = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int *idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    const uint8_t *p = data + (size_t)i * (size_t)cloud->point_step;
    x[i] = lfxo_load_f32(p + cloud->off_x);
    y[i] = lfxo_load_f32(p + cloud->off_y);
    ri[i].ring = lfxo_load_ring(p + cloud->off_ring, cloud->ring_datatype);
    ri[i].idx = i;
  }
  qsort(ri, (size_t)n, sizeof(lfxo_ri), lfxo_ri_cmp);

  for (int s = 0; s < n;) {
    int e = s;
    while (e < n && ri[e].ring == ri[s].ring) { e++; }
    const int m = e - s;
    if (m >= prm->padding + 1) { /* RemoveSparseRings(rings, padding+1) ring.cpp:46-59 */
      if (n_rings >= n_rings_cap) { rc = -1; break; }
      for (int k = 0; k < m; k++) { idx[k] = ri[s + k].idx; }
      lfxo_sort_by_polar_angle_f32(x, y, idx, m); /* SortEachRingByAngle ring.hpp:131-139 */
      for (int k = 0; k < m; k++) { rx[k] = x[idx[k]]; ry[k] = y[idx[k]]; sorted_src[n_sorted + k] = idx[k]; }
      const int skipped = lfxo_extract_ring(rx, ry, m, prm, labels + n_sorted, curvature + n_sorted);
      ring_ids[n_rings] = ri[s].ring;
      ring_sizes[n_rings] = m;
      ring_skipped[n_rings] = skipped;
      if (!skipped) { /* GetIndicesByValue lib/.../algorithm.hpp:50-62 */
        for (int k = 0; k < m; k++) {
          if (labels[n_sorted + k] == LFXO_EDGE) { edge_idx[n_edge++] = n_sorted + k; }
          if (labels[n_sorted + k] == LFXO_SURFACE) { surface_idx[n_surface++] = n_sorted + k; }
        }
      }
      n_sorted += m;
      n_rings++;
    }
    s = e;
  }
  free(ri); free(x); free(y); free(rx); free(ry); free(idx);
  *n_rings_out = n_rings;
  *n_edge_out = n_edge;
  *n_surface_out = n_surface;
  return rc < 0 ? rc : n_sorted;
}

/* ------------------------------------------------------------------ batch (CPU timing leg) */

typedef struct {
  const lfxo_cloud *clouds; int n_scans; const lfxo_params *prm; int tid, n_threads; int *counts; int rc;
} lfxo_job;

static void *lfxo_worker(void *arg)
{
  lfxo_job *job = (lfxo_job *)arg;
  for (int s = job->tid; s < job->n_scans; s += job->n_threads) {
    const int n = job->clouds[s].n_points, cap = n > 0 ? n : 1;
    int *ring_ids = (int *)malloc(sizeof(int) * (size_t)cap * 3);
    int *sorted_src = (int *)malloc(sizeof(int) * (size_t)cap * 3);
    uint8_t *labels = (uint8_t *)malloc((size_t)cap);
    double *curv = (double *)malloc(sizeof(double) * (size_t)cap);
    int n_rings, n_e, n_s;
    const int r = lfxo_extract_scan(&job->clouds[s], job->prm, cap, &n_rings, ring_ids, ring_ids + cap,
                                    ring_ids + 2 * cap, sorted_src, labels, curv,
                                    &n_e, sorted_src + cap, &n_s, sorted_src + 2 * cap);
    if (r < 0) { job->rc = r; }
    job->counts[2 * s] = n_e;
    job->counts[2 * s + 1] = n_s;
    free(ring_ids); free(sorted_src); free(labels); free(curv);
  }
  return NULL;
}

int lfxo_extract_batch_counts(const lfxo_cloud *clouds, int n_scans, const lfxo_params *prm,
                              int n_threads, int *counts)
{
  if (n_threads < 1) { n_threads = 1; }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  lfxo_job *jobs = (lfxo_job *)malloc(sizeof(lfxo_job) * (size_t)n_threads);
  int rc = 0;
  for (int t = 0; t < n_threads; t++) {
    jobs[t] = (lfxo_job){clouds, n_scans, prm, t, n_threads, counts, 0};
    pthread_create(&th[t], NULL, lfxo_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) { rc = jobs[t].rc; }
  }
  free(th); free(jobs);
  return rc;
}
