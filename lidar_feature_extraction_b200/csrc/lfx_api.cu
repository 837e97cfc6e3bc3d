// lfx_api.cu — host side of the C ABI declared in include/lfx.h.
//
// One handle = one GPU, one stream, grow-only device buffers, and a cache of CUDA graphs keyed by
// the batch geometry (number of scans, number of ingest tiles). Parameters are frozen at creation
// (the reference node holds `const HyperParameters params_`, feature_extraction.cpp:173); derived
// constants (the cosine cut replacing acos(c) < theta) are computed here on the host with the same
// libm the reference would use.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "lfx.h"
#include "lfx_kernels.cuh"
#include "lfx_ring.cuh"
#include "lfx_sector.cuh"
#include "lfx_synth.h"
#include "lfx_convert.cuh"
#include "lfx_color.cuh"
#include "lfx_map.cuh"
#include "lfx_shard.cuh"
#include "lfx_loc.cuh"
#include "lfx_big.cuh"

namespace lfxk
{
// lfx_sector_extra.cu, one translation unit per padding
void sector_kernels_p1(bool diag, void (**out)(const SectorArgs));
void sector_kernels_p3(bool diag, void (**out)(const SectorArgs));
void sector_kernels_p4(bool diag, void (**out)(const SectorArgs));
void sector_kernels_p6(bool diag, void (**out)(const SectorArgs));
void sector_kernels_p7(bool diag, void (**out)(const SectorArgs));
void sector_kernels_p8(bool diag, void (**out)(const SectorArgs));
}  // namespace lfxk

#include <dlfcn.h>
#include <nccl.h>      // types only: the library is loaded at run time (lfx_shard_*), single-GPU users do not need it
#include <sys/syscall.h>
#include <unistd.h>

using namespace lfxk;

namespace
{

thread_local std::string g_create_error;

template<typename T>
struct DevBuf
{
  T * p = nullptr;
  size_t cap = 0;  // elements
};

struct GraphKey
{
  int n_scans;
  uint32_t n_tiles;
  bool timed;   // the graph carries event-record nodes around the stages (lfx_set_stage_timing(h, 2))
  bool operator<(const GraphKey & o) const
  {
    if (n_scans != o.n_scans) { return n_scans < o.n_scans; }
    if (n_tiles != o.n_tiles) { return n_tiles < o.n_tiles; }
    return timed < o.timed;
  }
};

}  // namespace

struct lfx_handle
{
  lfx_params params{};
  lfx_options opt{};
  DevParams dev{};
  int device = 0;
  int num_sms = 0;
  int ring_grid = 0, pack_grid = 0, big_grid = 0;
  cudaStream_t aux_stream = nullptr;   // only ever captures the body of the batch graph's conditional node
  bool cond_ok = false;                // conditional graph nodes are used (LFX_NO_COND=1 turns them off)
  bool ring_enabled = true;   // the on-chip per-ring kernel covers these parameters (else every ring takes k_extract_rings_big)
  int tile = TILE_SMALL;   // points per ingest tile (TILE_BIG when the scatter's shared memory fits)
  size_t ring_smem = 0;
  int cap = 0;
  int ring_threads = 0;
  void (*ring_kernel)(const RingArgs) = nullptr;
  // [0..2]: regular scans (strided rings), [3..5]: indexed rings of bucketed scans; null: no fast path for these parameters
  void (*sector_kernel[2 * N_FAST_K])(const SectorArgs) = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int sector_grid[2 * N_FAST_K] = {0, 0, 0, 0, 0, 0}, ingest_grid = 0;
  size_t sector_smem[2 * N_FAST_K] = {0, 0, 0, 0, 0, 0};
  bool fast_enabled = false;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  uint64_t launches = 0;

  // device buffers
  DevBuf<ScanDesc> d_scans;
  DevBuf<uint64_t> d_point_base;
  DevBuf<uint16_t> d_ring16;
  DevBuf<uint32_t> d_tile_hist;
  DevBuf<lfx_ring_info> d_rings;
  DevBuf<uint2> d_work;
  DevBuf<uint2> d_ring_featoff;
  DevBuf<uint2> d_ring_src;
  DevBuf<uint32_t> d_scan_flags;
  DevBuf<uint32_t> d_gen_scan, d_gen_tile_base, d_tile_owner;
  DevBuf<FastRing> d_fast[2 * N_FAST_K];
  DevBuf<SectorRec> d_rec[2 * N_FAST_K];
  DevBuf<int> d_bndx[N_FAST_K];
  DevBuf<uint32_t> d_ring_path;
  DevBuf<uint32_t> d_idx;
  DevBuf<uint8_t> d_labels;
  DevBuf<uint32_t> d_sorted_src;
  DevBuf<double> d_curv;
  DevBuf<float4> d_stage, d_edge, d_surface;
  DevBuf<uint32_t> d_counts, d_offsets;
  DevBuf<uint8_t> d_input;
  uint32_t * d_counters = nullptr;
  // upstream converter (lfx_convert_batch)
  DevBuf<uint8_t> d_conv_raw, d_conv_out;
  DevBuf<uint16_t> d_conv_ring16;   // ring id of every converted point (k_convert by-product)
  DevBuf<ConvCloud> d_conv_clouds;
  DevBuf<unsigned long long> d_conv_state;
  DevBuf<uint32_t> d_conv_meta;     // ticket | kept[n] | flags[n]
  DevBuf<uint32_t> d_conv_tile_cloud;
  cudaEvent_t conv_ev[2] = {nullptr, nullptr};
  bool conv_smem_set = false;
  std::vector<uint64_t> conv_point_base;
  std::vector<uint32_t> conv_kept, conv_status;
  bool have_conv = false;
  // colored_scan (lfx_color_batch)
  DevBuf<uint4> d_colored;
  DevBuf<uint32_t> d_colored_counts;
  std::vector<uint32_t> colored_counts;
  std::vector<uint64_t> colored_base;
  bool have_colored = false;
  // mapping accumulate (lfx_map_add_batch)
  DevBuf<float4> d_map;
  DevBuf<MapFrame> d_map_frames;
  uint64_t map_points = 0;
  bool map_empty = true;
  double map_prev[12] = {0};

  // pinned host staging
  // descriptor tables: two pinned slots used in turn, so the host can lay out batch k+1 while batch k runs
  // (h_scans / h_point_base point at the slot of the current batch; desc_done[k] = slot k's H2D copies finished)
  ScanDesc * h_scans_buf = nullptr;
  uint64_t * h_point_base_buf = nullptr;
  ScanDesc * h_scans = nullptr;
  uint64_t * h_point_base = nullptr;
  size_t h_scans_cap = 0;
  cudaEvent_t desc_done[2] = {nullptr, nullptr};
  int desc_slot = 0;
  uint32_t * h_counters = nullptr;
  // single-scan convenience mirrors
  float4 * h_edge = nullptr, * h_surface = nullptr;
  uint8_t * h_labels = nullptr;
  uint32_t * h_sorted_src = nullptr;
  size_t h_scan_cap = 0;

  std::map<GraphKey, cudaGraphExec_t> graphs;

  // last batch
  int n_scans = 0;
  uint64_t total_points = 0;
  uint32_t total_tiles = 0;
  bool have_batch = false;

  struct LocState * loc = nullptr;   // localization residual build (lfx_loc_*), created on first use

  bool timing = false;
  bool timing_in_graph = false;   // stage events as external event-record nodes of the captured graph
  cudaEvent_t ev[LFX_N_STAGES + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool have_timing = false;
};

namespace
{

#define LFX_CUDA(h, call)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                              \
      return LFX_E_CUDA;                                                                           \
    }                                                                                              \
  } while (0)

int fail(lfx_handle * h, int code, const std::string & msg)
{
  h->err = msg;
  return code;
}

template<typename T>
int ensure(lfx_handle * h, DevBuf<T> & b, size_t n, bool * regrown)
{
  if (n <= b.cap) { return LFX_OK; }
  // all work using the old buffer must be finished before it is released
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (b.p) { LFX_CUDA(h, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
  const size_t want = n + n / 8 + 64;
  LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&b.p), want * sizeof(T)));
  b.cap = want;
  if (regrown) { *regrown = true; }
  return LFX_OK;
}

void drop_graphs(lfx_handle * h)
{
  for (auto & kv : h->graphs) { cudaGraphExecDestroy(kv.second); }
  h->graphs.clear();
}

// Smallest double c with acos(c) < theta; is_neighbor <=> c_min <= c <= 1.
// acos is the host libm's, i.e. the one the reference itself calls (math.cpp:45).
double cos_cut(double theta)
{
  auto key = [](double v) { int64_t b; memcpy(&b, &v, 8); return b < 0 ? (int64_t)(0x8000000000000000ull - (uint64_t)b) : b; };
  auto unkey = [](int64_t k) { int64_t b = k < 0 ? (int64_t)(0x8000000000000000ull - (uint64_t)k) : k; double v; memcpy(&v, &b, 8); return v; };
  if (!(std::acos(1.0) < theta)) { return 2.0; }       // nothing is ever a neighbour
  if (std::acos(-1.0) < theta) { return -1.0; }        // everything with a finite cosine is
  int64_t lo = key(-1.0), hi = key(1.0);               // pred(lo) false, pred(hi) true
  while (hi - lo > 1) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (std::acos(unkey(mid)) < theta) { hi = mid; } else { lo = mid; }
  }
  return unkey(hi);
}

// Smallest double q with (double)(float)q > rho: the parallel-beam test of parallel_beam.hpp:44-47
// narrows the ratio to float before comparing, so the predicate is monotone in q and has a single cut.
double ratio_cut(double rho)
{
  auto key = [](double v) { int64_t b; memcpy(&b, &v, 8); return b; };          // q >= 0 only
  auto unkey = [](int64_t k) { double v; memcpy(&v, &k, 8); return v; };
  auto pred = [&](double q) { const float f = (float)q; return (double)f > rho; };
  int64_t lo = key(0.0), hi = key(1.7976931348623157e308);  // pred(lo) false (rho > 0), pred(hi) true ((float)DBL_MAX = inf)
  if (!pred(unkey(hi))) { return INFINITY; }
  while (hi - lo > 1) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (pred(unkey(mid))) { hi = mid; } else { lo = mid; }
  }
  return unkey(hi);
}

template<int PT>
void (*pick_ring_kernel_t(int threads, int * tmax))(const RingArgs)
{
  if (threads <= 288) { *tmax = 288; return k_extract_rings<PT, 288, 2>; }
  if (threads <= 512) { *tmax = 512; return k_extract_rings<PT, 512, 1>; }
  *tmax = 1024;
  return k_extract_rings<PT, 1024, 1>;
}

void (*pick_ring_kernel(int padding, int threads, int * tmax))(const RingArgs)
{
  if (padding == 5) { return pick_ring_kernel_t<5>(threads, tmax); }   // compiled default
  if (padding == 2) { return pick_ring_kernel_t<2>(threads, tmax); }   // launch YAML
  return pick_ring_kernel_t<0>(threads, tmax);                         // any other padding: runtime loops
}

template<int P, bool DIAG>
void pick_sector_kernels_t(void (**out)(const SectorArgs))
{
  out[0] = k_extract_sectors<P, fast_k(0), DIAG, false>;
  out[1] = k_extract_sectors<P, fast_k(1), DIAG, false>;
  out[2] = k_extract_sectors<P, fast_k(2), DIAG, false>;
  out[3] = k_extract_sectors<P, fast_k(0), DIAG, true>;
  out[4] = k_extract_sectors<P, fast_k(1), DIAG, true>;
  out[5] = k_extract_sectors<P, fast_k(2), DIAG, true>;
}

// The sector kernel is compiled for convolution_padding 1..8: the two deployed paddings (compiled default 5, launch
// YAML 2) here, the others in their own translation units (lfx_sector_extra.cu); any other padding runs on the per-ring
// kernels only.
bool pick_sector_kernels(int padding, bool diag, void (**out)(const SectorArgs))
{
  if (padding == 5) { if (diag) { pick_sector_kernels_t<5, true>(out); } else { pick_sector_kernels_t<5, false>(out); } return true; }
  if (padding == 2) { if (diag) { pick_sector_kernels_t<2, true>(out); } else { pick_sector_kernels_t<2, false>(out); } return true; }
#ifdef LFX_EXTRA_PADDINGS
  switch (padding) {
    case 1: lfxk::sector_kernels_p1(diag, out); return true;
    case 3: lfxk::sector_kernels_p3(diag, out); return true;
    case 4: lfxk::sector_kernels_p4(diag, out); return true;
    case 6: lfxk::sector_kernels_p6(diag, out); return true;
    case 7: lfxk::sector_kernels_p7(diag, out); return true;
    case 8: lfxk::sector_kernels_p8(diag, out); return true;
    default: break;
  }
#endif
  return false;
}

int validate_params(const lfx_params & p, std::string & why)
{
  // hyper_parameter.hpp:45-53: everything strictly positive
  if (!(p.padding > 0) || !(p.neighbor_degree_threshold > 0) || !(p.distance_diff_threshold > 0) ||
      !(p.parallel_beam_min_range_ratio > 0) || !(p.edge_threshold > 0) || !(p.surface_threshold > 0) ||
      !(p.min_range > 0) || !(p.max_range > 0) || !(p.n_blocks > 0)) {
    why = "all nine parameters must be > 0 (hyper_parameter.hpp:45-53)";
    return LFX_E_BAD_PARAM;
  }
  // no upper bounds, like the reference: paddings above MAX_PADDING and more than MAX_BLOCKS sectors run on
  // k_extract_rings_big (lfx_big.cuh); only the arithmetic on positions has to stay inside 32 bits
  if (p.padding > (1 << 24) || p.n_blocks > (1 << 24)) { why = "convolution_padding / n_blocks above 2^24"; return LFX_E_BAD_PARAM; }
  return LFX_OK;
}

constexpr int COND_MIN_SCANS = 32;   // smallest batch whose graph wraps the general path in a conditional node

// ---- the batch pipeline (enqueued on h->stream; also the body of the captured graph)
int enqueue_pipeline(lfx_handle * h, int n_scans, uint32_t n_tiles, bool with_events, bool capturing = false)
{
  // inside a stream capture a plain cudaEventRecord only marks a dependency; the External flag makes it a real
  // event-record node that is timestamped on every launch of the graph
  const unsigned ev_flags = capturing ? cudaEventRecordExternal : cudaEventRecordDefault;
  const int max_rings = h->opt.max_rings;
  LFX_CUDA(h, cudaMemsetAsync(h->d_counters, 0, sizeof(uint32_t) * C_COUNT, h->stream));
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[0], h->stream, ev_flags)); }
  // ---- fast path: layout probe, then one warp per ring-sector
  ProbeArgs pa;
  pa.scans = h->d_scans.p;
  pa.rings = h->d_rings.p;
  pa.scan_flags = h->d_scan_flags.p;
  for (int c = 0; c < N_FAST_K; c++) { pa.fast[c] = h->d_fast[c].p; }
  pa.counters = h->d_counters;
  pa.max_rings = max_rings;
  pa.P = h->params.padding;
  pa.B = h->params.n_blocks;
  pa.enabled = h->fast_enabled ? 1 : 0;
  k_probe_layout<<<n_scans, PROBE_LAYOUT_THREADS, sizeof(uint32_t) * (3 * max_rings + 1), h->stream>>>(pa);
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[1], h->stream, ev_flags)); }
  if (h->fast_enabled) {
    for (int c = 0; c < N_FAST_K; c++) {
      SectorArgs sa;
      sa.fast = h->d_fast[c].p;
      sa.bnd = nullptr;
      sa.n_entries = h->d_counters + C_N_FAST0 + c;
      sa.ring_path = nullptr;
      sa.work = nullptr;
      sa.counters = h->d_counters;
      sa.rec = h->d_rec[c].p;
      sa.rings = h->d_rings.p;
      sa.scan_flags = h->d_scan_flags.p;
      sa.labels = h->d_labels.p;
      sa.sorted_src = h->opt.want_sorted_src ? h->d_sorted_src.p : nullptr;
      sa.curvature = h->opt.want_curvature ? h->d_curv.p : nullptr;
      sa.stage = h->d_stage.p;
      sa.max_rings = max_rings;
      sa.inv_blocks = (uint32_t)(0x100000000ull / (uint64_t)h->params.n_blocks);
      sa.prm = h->dev;
      h->sector_kernel[c]<<<h->sector_grid[c], sector_warps(fast_k(c), false) * 32, h->sector_smem[c], h->stream>>>(sa);
    }
  }
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[2], h->stream, ev_flags)); }
  // ---- general path for the scans flagged by the probe or by a failed check of the sector kernel
  // Inside a captured graph (without stage events) everything between here and the packing is the body of an IF node
  // whose condition k_general_list sets: a batch of regular scans skips the nine launches of the general path.
  cudaGraphConditionalHandle cond_handle = 0;
  cudaGraph_t cap_graph = nullptr;
  // (only for batches of at least COND_MIN_SCANS scans: for a handful of scans the node costs about what it saves, and
  //  the per-scan entry of the ROS callback stays a plain sequence of kernel nodes)
  const bool use_cond = capturing && !with_events && h->cond_ok && n_scans >= COND_MIN_SCANS;
  if (use_cond) {
    cudaStreamCaptureStatus cs;
    LFX_CUDA(h, cudaStreamGetCaptureInfo(h->stream, &cs, nullptr, &cap_graph, nullptr, nullptr));
    LFX_CUDA(h, cudaGraphConditionalHandleCreate(&cond_handle, cap_graph, 0, cudaGraphCondAssignDefault));
  }
  k_general_list<<<1, 1024, 0, h->stream>>>(h->d_scans.p, n_scans, h->d_scan_flags.p, h->d_gen_scan.p, h->d_gen_tile_base.p,
                                            h->d_tile_owner.p, h->d_counters, cond_handle, use_cond ? 1 : 0);
  struct StreamRestore   // an error return below must not leave the handle on the capture stream of the body
  {
    lfx_handle * h; cudaStream_t s;
    ~StreamRestore() { h->stream = s; }
  } restore{h, h->stream};
  cudaStream_t outer_stream = h->stream;
  if (use_cond) {
    cudaStreamCaptureStatus cs;
    const cudaGraphNode_t * deps = nullptr;
    size_t n_deps = 0;
    LFX_CUDA(h, cudaStreamGetCaptureInfo(h->stream, &cs, nullptr, &cap_graph, &deps, &n_deps));
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = cond_handle;
    np.conditional.type = cudaGraphCondTypeIf;
    np.conditional.size = 1;
    cudaGraphNode_t cond_node = nullptr;
    LFX_CUDA(h, cudaGraphAddNode(&cond_node, cap_graph, deps, n_deps, &np));
    LFX_CUDA(h, cudaStreamUpdateCaptureDependencies(h->stream, &cond_node, 1, cudaStreamSetCaptureDependencies));
    LFX_CUDA(h, cudaStreamBeginCaptureToGraph(h->aux_stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    h->stream = h->aux_stream;   // the launches below go into the body
  }
  const int ingest_grid = (int)std::min<uint32_t>(std::max<uint32_t>(n_tiles, 1u), (uint32_t)h->ingest_grid);
  if (h->tile == TILE_BIG) {
    k_ring_hist<TILE_BIG><<<ingest_grid, INGEST_THREADS, sizeof(uint32_t) * max_rings, h->stream>>>(
      h->d_scans.p, h->d_gen_scan.p, h->d_gen_tile_base.p, h->d_tile_owner.p, h->d_ring16.p, h->d_tile_hist.p, max_rings, h->d_counters);
  } else {
    k_ring_hist<TILE_SMALL><<<ingest_grid, INGEST_THREADS, sizeof(uint32_t) * max_rings, h->stream>>>(
      h->d_scans.p, h->d_gen_scan.p, h->d_gen_tile_base.p, h->d_tile_owner.p, h->d_ring16.p, h->d_tile_hist.p, max_rings, h->d_counters);
  }
  k_ring_plan<<<n_scans, 256, sizeof(uint32_t) * 2 * max_rings, h->stream>>>(
    h->d_scans.p, h->d_scan_flags.p, h->d_tile_hist.p, h->d_rings.p, h->d_ring_src.p, max_rings, h->params.padding,
    h->ring_enabled ? h->cap : 0);   // longer rings are marked LFX_RING_TOO_LONG: the big kernel's share
  if (h->tile == TILE_BIG) {
    k_ring_scatter<TILE_BIG><<<ingest_grid, INGEST_THREADS, scatter_smem_bytes(max_rings, TILE_BIG), h->stream>>>(
      h->d_scans.p, h->d_gen_scan.p, h->d_gen_tile_base.p, h->d_tile_owner.p, h->d_counters, h->d_ring16.p, h->d_tile_hist.p, h->d_rings.p, h->d_idx.p, max_rings);
  } else {
    k_ring_scatter<TILE_SMALL><<<ingest_grid, INGEST_THREADS, scatter_smem_bytes(max_rings, TILE_SMALL), h->stream>>>(
      h->d_scans.p, h->d_gen_scan.p, h->d_gen_tile_base.p, h->d_tile_owner.p, h->d_counters, h->d_ring16.p, h->d_tile_hist.p, h->d_rings.p, h->d_idx.p, max_rings);
  }
  // ---- bucketed rings that are rotated monotone sequences: sector kernel through the index list; the rest
  //      (and every ring whose hypothesis fails there) form the work list of the per-ring kernel
  RingProbeArgs rp;
  rp.scans = h->d_scans.p;
  rp.gen_scan = h->d_gen_scan.p;
  rp.idx = h->d_idx.p;
  rp.rings = h->d_rings.p;
  for (int c = 0; c < N_FAST_K; c++) { rp.fastx[c] = h->d_fast[N_FAST_K + c].p; rp.bndx[c] = h->d_bndx[c].p; }
  rp.ring_path = h->d_ring_path.p;
  rp.work = h->d_work.p;
  rp.counters = h->d_counters;
  rp.max_rings = max_rings;
  rp.P = h->params.padding;
  rp.B = h->params.n_blocks;
  rp.enabled = h->fast_enabled ? 1 : 0;
  k_probe_rings<<<n_scans, PROBE_THREADS, sizeof(uint32_t) * 3 * max_rings, h->stream>>>(rp);
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[3], h->stream, ev_flags)); }
  if (h->fast_enabled) {
    for (int c = 0; c < N_FAST_K; c++) {
      SectorArgs sa;
      sa.fast = h->d_fast[N_FAST_K + c].p;
      sa.bnd = h->d_bndx[c].p;
      sa.n_entries = h->d_counters + C_N_FASTX0 + c;
      sa.ring_path = h->d_ring_path.p;
      sa.work = h->d_work.p;
      sa.counters = h->d_counters;
      sa.rec = h->d_rec[N_FAST_K + c].p;
      sa.rings = h->d_rings.p;
      sa.scan_flags = h->d_scan_flags.p;
      sa.labels = h->d_labels.p;
      sa.sorted_src = h->opt.want_sorted_src ? h->d_sorted_src.p : nullptr;
      sa.curvature = h->opt.want_curvature ? h->d_curv.p : nullptr;
      sa.stage = h->d_stage.p;
      sa.max_rings = max_rings;
      sa.inv_blocks = (uint32_t)(0x100000000ull / (uint64_t)h->params.n_blocks);
      sa.prm = h->dev;
      h->sector_kernel[N_FAST_K + c]<<<h->sector_grid[N_FAST_K + c], sector_warps(fast_k(c), true) * 32, h->sector_smem[N_FAST_K + c], h->stream>>>(sa);
    }
  }
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[4], h->stream, ev_flags)); }
  RingArgs ra;
  ra.scans = h->d_scans.p;
  ra.idx = h->d_idx.p;
  ra.rings = h->d_rings.p;
  ra.ring_src = h->d_ring_src.p;
  ra.scan_flags = h->d_scan_flags.p;
  ra.work = h->d_work.p;
  ra.counters = h->d_counters;
  ra.labels = h->d_labels.p;
  ra.sorted_src = h->opt.want_sorted_src ? h->d_sorted_src.p : nullptr;
  ra.curvature = h->opt.want_curvature ? h->d_curv.p : nullptr;
  ra.stage = h->d_stage.p;
  ra.max_rings = max_rings;
  ra.cap = h->cap;
  ra.force_order_path = h->opt.force_order_path;
  ra.prm = h->dev;
  if (h->ring_enabled) { h->ring_kernel<<<h->ring_grid, h->ring_threads, h->ring_smem, h->stream>>>(ra); }
  {
    BigArgs ba;
    ba.scans = h->d_scans.p;
    ba.idx = h->d_idx.p;
    ba.rings = h->d_rings.p;
    ba.work = h->d_work.p;
    ba.counters = h->d_counters;
    ba.labels = h->d_labels.p;
    ba.sorted_src = ra.sorted_src;
    ba.curvature = ra.curvature;
    ba.stage = h->d_stage.p;
    ba.tmp16 = h->d_ring16.p;
    ba.max_rings = max_rings;
    ba.all = h->ring_enabled ? 0 : 1;
    ba.prm = h->dev;
    k_extract_rings_big<<<h->big_grid, BIG_THREADS, 0, h->stream>>>(ba);
  }
  if (use_cond) {
    h->stream = outer_stream;
    cudaGraph_t body = nullptr;
    LFX_CUDA(h, cudaStreamEndCapture(h->aux_stream, &body));
  }
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[5], h->stream, ev_flags)); }
  // ---- packing
  k_feat_offsets_a<<<n_scans, 128, 0, h->stream>>>(h->d_rings.p, h->d_ring_featoff.p, h->d_counts.p, max_rings);
  k_feat_offsets_b<<<1, 1024, 0, h->stream>>>(h->d_counts.p, h->d_offsets.p, n_scans);
  if (h->fast_enabled) {
    PackFastArgs pf;
    for (int c = 0; c < 2 * N_FAST_K; c++) { pf.fast[c] = h->d_fast[c].p; pf.rec[c] = h->d_rec[c].p; }
    pf.counters = h->d_counters;
    pf.scan_flags = h->d_scan_flags.p;
    pf.ring_path = h->d_ring_path.p;
    pf.ring_featoff = h->d_ring_featoff.p;
    pf.offsets = h->d_offsets.p;
    pf.stage = h->d_stage.p;
    pf.edge = h->d_edge.p;
    pf.surface = h->d_surface.p;
    pf.max_rings = max_rings;
    pf.B = h->params.n_blocks;
    k_pack_fast<<<h->pack_grid, 256, 0, h->stream>>>(pf);
  }
  k_pack_copy<<<h->pack_grid, 256, 0, h->stream>>>(
    h->d_work.p, h->d_counters, h->d_scans.p, h->d_rings.p, h->d_ring_featoff.p, h->d_offsets.p, h->d_stage.p,
    h->d_edge.p, h->d_surface.p, max_rings);
  if (with_events) { LFX_CUDA(h, cudaEventRecordWithFlags(h->ev[6], h->stream, ev_flags)); }
  LFX_CUDA(h, cudaGetLastError());
  return LFX_OK;
}

// kernels of one batch; `conditional`: the batch ran as a graph whose general path is the body of an IF node - its
// kernels (ingest x3, ring probe, indexed sector kernels, per-ring kernels) are then not counted: a lower bound
int kernels_per_batch(const lfx_handle * h, bool conditional)
{
  const int all = (h->fast_enabled ? 11 + 2 * N_FAST_K + 1 : 11) + (h->ring_enabled ? 1 : 0);
  const int body = 4 + (h->fast_enabled ? N_FAST_K : 0) + (h->ring_enabled ? 1 : 0) + 1;
  return conditional ? all - body : all;
}

}  // namespace

// ====================================================================== C ABI

extern "C" {

void lfx_default_params(lfx_params * out)
{
  // hyper_parameter.hpp:35-43
  out->padding = 5;
  out->neighbor_degree_threshold = 2.0;
  out->distance_diff_threshold = 0.3;
  out->parallel_beam_min_range_ratio = 0.02;
  out->edge_threshold = 0.05;
  out->surface_threshold = 0.05;
  out->min_range = 0.1;
  out->max_range = 100.0;
  out->n_blocks = 6;
}

void lfx_launch_yaml_params(lfx_params * out)
{
  // lidar_feature_launch/config/lidar_feature_extraction.param.yaml:3-10; surface_threshold is not
  // in the YAML and keeps its compiled default
  lfx_default_params(out);
  out->padding = 2;
  out->neighbor_degree_threshold = 3.0;
  out->distance_diff_threshold = 0.3;
  out->parallel_beam_min_range_ratio = 0.02;
  out->edge_threshold = 50.0;
  out->min_range = 0.1;
  out->max_range = 1000.0;
  out->n_blocks = 6;
}

const char * lfx_last_error(const lfx_handle * h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lfx_create(const lfx_params * params, const lfx_options * options, lfx_handle ** out)
{
  if (!params || !out) { g_create_error = "null argument"; return LFX_E_BAD_PARAM; }
  *out = nullptr;
  std::string why;
  if (int rc = validate_params(*params, why)) { g_create_error = why; return rc; }
  lfx_handle * h = new lfx_handle();
  h->params = *params;
  if (options) { h->opt = *options; }
  // default: a 4096-column sweep / an HDL-64 at 5 Hz still fits the on-chip per-ring kernel (174 KB of shared memory at
  // padding 5); regular and rotated-monotone rings of up to 372-point sectors never get there (sector kernel)
  if (h->opt.max_ring_points <= 0) { h->opt.max_ring_points = 4096; }
  if (h->opt.max_rings <= 0) { h->opt.max_rings = 128; }
  if (h->opt.max_rings > 4096 || h->opt.force_order_path < 0 || h->opt.force_order_path > 2) {
    g_create_error = "lfx_options outside the supported envelope (max_rings <= 4096)";
    delete h;
    return LFX_E_BAD_PARAM;
  }
  h->device = h->opt.device;
  // The on-chip per-ring kernel holds rings of up to `cap` points (at most 8192: 1024 threads x 8 positions) and is
  // compiled for paddings <= MAX_PADDING and <= MAX_BLOCKS sectors; whatever lies beyond runs on k_extract_rings_big.
  h->ring_enabled = params->padding <= MAX_PADDING && params->n_blocks <= MAX_BLOCKS;
  h->opt.max_ring_points = std::min(h->opt.max_ring_points, 8192);
  h->cap = (std::max(h->opt.max_ring_points, h->ring_enabled ? 2 * params->padding + 2 : 0) + 255) & ~255;
  h->ring_threads = h->cap / PTS;

  auto bail = [&](cudaError_t e, const char * what) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
    lfx_destroy(h);
    return LFX_E_CUDA;
  };
  cudaError_t e;
  int n_dev = 0;
  if ((e = cudaGetDeviceCount(&n_dev)) != cudaSuccess || n_dev == 0) {
    g_create_error = std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
    delete h;
    return LFX_E_CUDA;
  }
  if ((e = cudaSetDevice(h->device)) != cudaSuccess) { return bail(e, "cudaSetDevice"); }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) { return bail(e, "cudaGetDeviceProperties"); }
  h->num_sms = prop.multiProcessorCount;
  if (h->opt.stream) { h->stream = reinterpret_cast<cudaStream_t>(h->opt.stream); }
  else {
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) { return bail(e, "cudaStreamCreate"); }
    h->own_stream = true;
  }

  if ((e = cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking)) != cudaSuccess) { return bail(e, "cudaStreamCreate(aux)"); }
  {
    const char * nc = getenv("LFX_NO_COND");
    h->cond_ok = !(nc && atoi(nc) != 0);
  }

  // derived constants
  h->dev.P = params->padding;
  h->dev.B = params->n_blocks;
  h->dev.c_min = cos_cut(params->neighbor_degree_threshold * M_PI / 180.0);  // DegreeToRadian degree_to_radian.hpp:34-37
  h->dev.c_lo = h->dev.c_min * (1.0 - 0x1p-40);
  h->dev.c_hi = h->dev.c_min * (1.0 + 0x1p-40);
  h->dev.q_min = ratio_cut(params->parallel_beam_min_range_ratio);
  h->dev.q_lo = h->dev.q_min * (1.0 - 0x1p-40);
  h->dev.q_hi = h->dev.q_min * (1.0 + 0x1p-40);
  h->dev.d = params->distance_diff_threshold;
  h->dev.rho = params->parallel_beam_min_range_ratio;
  h->dev.tau_e = params->edge_threshold;
  h->dev.tau_s = params->surface_threshold;
  h->dev.rmin = params->min_range;
  h->dev.rmax = params->max_range;
  h->dev.center_w = -2. * params->padding;  // MakeWeight curvature.cpp:40

  // kernel attributes / persistent grid sizes
  int occ = 0;
  if (h->ring_enabled) {
    // (a capacity that does not fit the shared memory of this device is lowered: longer rings take the big kernel)
    while (h->cap > 256 && ring_smem_bytes(h->cap, params->padding) > (size_t)prop.sharedMemPerBlockOptin) { h->cap -= 256; }
    h->ring_threads = h->cap / PTS;
    h->ring_smem = ring_smem_bytes(h->cap, params->padding);
    if (h->ring_smem > (size_t)prop.sharedMemPerBlockOptin) { h->ring_enabled = false; }
  }
  if (h->ring_enabled) {
    int tmax = 0;
    h->ring_kernel = pick_ring_kernel(params->padding, h->ring_threads, &tmax);
    if ((e = cudaFuncSetAttribute(h->ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ring_smem)) != cudaSuccess) { return bail(e, "cudaFuncSetAttribute(rings)"); }
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->ring_kernel, h->ring_threads, h->ring_smem)) != cudaSuccess) { return bail(e, "occupancy(rings)"); }
    h->ring_grid = h->num_sms * std::max(occ, 1);
  }
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_extract_rings_big, BIG_THREADS, 0)) != cudaSuccess) { return bail(e, "occupancy(big rings)"); }
  h->big_grid = h->num_sms * std::max(occ, 1);
  // ingest tile: the big one while three scatter CTAs still fit an SM
  h->tile = scatter_smem_bytes(h->opt.max_rings, TILE_BIG) * 3 <= (size_t)prop.sharedMemPerMultiprocessor ? TILE_BIG : TILE_SMALL;
  const size_t scatter_smem = scatter_smem_bytes(h->opt.max_rings, h->tile);
  if (scatter_smem > 48 * 1024) {
    e = h->tile == TILE_BIG ? cudaFuncSetAttribute(k_ring_scatter<TILE_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem)
                            : cudaFuncSetAttribute(k_ring_scatter<TILE_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem);
    if (e != cudaSuccess) { return bail(e, "cudaFuncSetAttribute(scatter)"); }
  }
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pack_copy, 256, 0)) != cudaSuccess) { return bail(e, "occupancy(pack)"); }
  h->pack_grid = h->num_sms * std::max(occ, 1);
  h->ingest_grid = h->num_sms * 8;
  // fast path: compiled for the deployed paddings; the sort-path test knob forces the general kernel
  h->fast_enabled = h->opt.force_order_path == 0 && params->n_blocks <= FAST_MAX_BLOCKS &&
                    pick_sector_kernels(params->padding, h->opt.want_sorted_src || h->opt.want_curvature, h->sector_kernel);
  if (h->fast_enabled) {
    h->sector_smem[0] = sector_smem_bytes<fast_k(0), false>(sector_warps(fast_k(0), false));
    h->sector_smem[1] = sector_smem_bytes<fast_k(1), false>(sector_warps(fast_k(1), false));
    h->sector_smem[2] = sector_smem_bytes<fast_k(2), false>(sector_warps(fast_k(2), false));
    h->sector_smem[3] = sector_smem_bytes<fast_k(0), true>(sector_warps(fast_k(0), true));
    h->sector_smem[4] = sector_smem_bytes<fast_k(1), true>(sector_warps(fast_k(1), true));
    h->sector_smem[5] = sector_smem_bytes<fast_k(2), true>(sector_warps(fast_k(2), true));
    for (int c = 0; c < 2 * N_FAST_K; c++) {
      if ((e = cudaFuncSetAttribute(h->sector_kernel[c], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->sector_smem[c])) != cudaSuccess) { return bail(e, "cudaFuncSetAttribute(sectors)"); }
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->sector_kernel[c], sector_warps(fast_k(c % N_FAST_K), c >= N_FAST_K) * 32, h->sector_smem[c])) != cudaSuccess) { return bail(e, "occupancy(sectors)"); }
      if (occ < 1) { g_create_error = "sector kernel does not fit on this device"; lfx_destroy(h); return LFX_E_CUDA; }
      // LFX_RESERVE_SMS (experiments): leave that many SMs to kernels of other streams (an overlapped collective)
      const char * rs = getenv("LFX_RESERVE_SMS");
      const int reserve = rs ? std::max(0, std::min(atoi(rs), h->num_sms - 1)) : 0;
      h->sector_grid[c] = (h->num_sms - reserve) * occ;
    }
  }
  const size_t probe_smem = sizeof(uint32_t) * (3 * (size_t)h->opt.max_rings + 1);
  if (probe_smem > 48 * 1024) {
    if ((e = cudaFuncSetAttribute(k_probe_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem)) != cudaSuccess) { return bail(e, "cudaFuncSetAttribute(probe)"); }
    if ((e = cudaFuncSetAttribute(k_probe_rings, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe_smem)) != cudaSuccess) { return bail(e, "cudaFuncSetAttribute(ring probe)"); }
  }

  if ((e = cudaMalloc(reinterpret_cast<void **>(&h->d_counters), sizeof(uint32_t) * C_COUNT)) != cudaSuccess) { return bail(e, "cudaMalloc(counters)"); }
  if ((e = cudaMallocHost(reinterpret_cast<void **>(&h->h_counters), sizeof(uint32_t) * C_COUNT)) != cudaSuccess) { return bail(e, "cudaMallocHost(counters)"); }
  for (auto & ev : h->ev) {
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) { return bail(e, "cudaEventCreate"); }
  }
  *out = h;
  return LFX_OK;
}

extern "C" int lfx_loc_release(lfx_handle * h);

void lfx_destroy(lfx_handle * h)
{
  if (!h) { return; }
  lfx_loc_release(h);
  cudaSetDevice(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); }
  drop_graphs(h);
  cudaFree(h->d_scans.p); cudaFree(h->d_point_base.p); cudaFree(h->d_ring16.p); cudaFree(h->d_tile_hist.p);
  cudaFree(h->d_rings.p); cudaFree(h->d_work.p); cudaFree(h->d_ring_featoff.p); cudaFree(h->d_ring_src.p); cudaFree(h->d_scan_flags.p); cudaFree(h->d_idx.p);
  cudaFree(h->d_labels.p); cudaFree(h->d_sorted_src.p); cudaFree(h->d_curv.p); cudaFree(h->d_stage.p);
  cudaFree(h->d_edge.p); cudaFree(h->d_surface.p); cudaFree(h->d_counts.p); cudaFree(h->d_offsets.p);
  cudaFree(h->d_input.p); cudaFree(h->d_counters); cudaFree(h->d_gen_scan.p); cudaFree(h->d_gen_tile_base.p); cudaFree(h->d_tile_owner.p);
  for (int c = 0; c < 2 * N_FAST_K; c++) { cudaFree(h->d_fast[c].p); cudaFree(h->d_rec[c].p); }
  for (int c = 0; c < N_FAST_K; c++) { cudaFree(h->d_bndx[c].p); }
  cudaFree(h->d_ring_path.p);
  cudaFree(h->d_conv_raw.p); cudaFree(h->d_conv_out.p); cudaFree(h->d_conv_ring16.p); cudaFree(h->d_conv_clouds.p); cudaFree(h->d_conv_state.p); cudaFree(h->d_conv_meta.p); cudaFree(h->d_conv_tile_cloud.p);
  cudaFree(h->d_colored.p); cudaFree(h->d_colored_counts.p);
  cudaFree(h->d_map.p); cudaFree(h->d_map_frames.p);
  for (auto & ev : h->conv_ev) { if (ev) { cudaEventDestroy(ev); } }
  cudaFreeHost(h->h_scans_buf); cudaFreeHost(h->h_point_base_buf); cudaFreeHost(h->h_counters);
  for (auto & ev : h->desc_done) { if (ev) { cudaEventDestroy(ev); } }
  cudaFreeHost(h->h_edge); cudaFreeHost(h->h_surface); cudaFreeHost(h->h_labels); cudaFreeHost(h->h_sorted_src);
  for (auto & ev : h->ev) { if (ev) { cudaEventDestroy(ev); } }
  if (h->aux_stream) { cudaStreamDestroy(h->aux_stream); }
  if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); }
  delete h;
}

int lfx_get_params(const lfx_handle * h, lfx_params * out)
{
  if (!h || !out) { return LFX_E_BAD_PARAM; }
  *out = h->params;
  return LFX_OK;
}

int lfx_device(const lfx_handle * h) { return h ? h->device : -1; }

void * lfx_stream(const lfx_handle * h) { return h ? reinterpret_cast<void *>(h->stream) : nullptr; }

int lfx_extract_batch(lfx_handle * h, const lfx_cloud_view * scans, int n_scans, lfx_batch_result * out)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  h->have_colored = false;
  if (n_scans < 0 || (n_scans > 0 && !scans)) { return fail(h, LFX_E_BAD_PARAM, "bad scans argument"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  h->have_batch = false;
  h->have_timing = false;

  // ---- validate (feature_extraction.cpp:96-108) and lay the batch out
  uint64_t total_points = 0, total_tiles = 0;
  size_t input_bytes = 0;
  for (int s = 0; s < n_scans; s++) {
    const lfx_cloud_view & v = scans[s];
    if (!v.is_dense) { return fail(h, LFX_E_NOT_DENSE, "Point cloud is not in dense format, please remove NaN points first!"); }
    if (!v.has_ring) { return fail(h, LFX_E_NO_RING, "Ring channel could not be found"); }
    if (v.ring_datatype != LFX_RING_U8 && v.ring_datatype != LFX_RING_U16 && v.ring_datatype != LFX_RING_U32) {
      return fail(h, LFX_E_BAD_LAYOUT, "ring field must be UINT8, UINT16 or UINT32");
    }
    const uint32_t ring_bytes = v.ring_datatype == LFX_RING_U8 ? 1 : (v.ring_datatype == LFX_RING_U16 ? 2 : 4);
    if (v.n_points > 0 && !v.data) { return fail(h, LFX_E_BAD_LAYOUT, "null data with n_points > 0"); }
    // (64-bit sums: offsets near 2^32 must not wrap past the check)
    if (v.point_step < 12 || (uint64_t)v.off_x + 4 > v.point_step || (uint64_t)v.off_y + 4 > v.point_step ||
        (uint64_t)v.off_z + 4 > v.point_step || (uint64_t)v.off_ring + ring_bytes > v.point_step) {
      return fail(h, LFX_E_BAD_LAYOUT, "field offsets exceed point_step");
    }
    if ((v.point_step | v.off_x | v.off_y | v.off_z) % 4 != 0 || v.off_ring % ring_bytes != 0 ||
        (v.memory == LFX_MEM_DEVICE && reinterpret_cast<uintptr_t>(v.data) % 4 != 0)) {
      return fail(h, LFX_E_BAD_LAYOUT, "x/y/z must be 4-byte aligned and ring naturally aligned");
    }
    total_points += v.n_points;
    total_tiles += (v.n_points + h->tile - 1) / h->tile;
    if (v.memory == LFX_MEM_HOST) { input_bytes += ((size_t)v.n_points * v.point_step + 15) & ~(size_t)15; }
  }
  if (total_points >= 0xFFFF0000ull) { return fail(h, LFX_E_CAPACITY, "batch exceeds 2^32 points"); }

  // ---- capacity
  const size_t np = std::max<uint64_t>(total_points, 1), ns = std::max(n_scans, 1);
  const size_t mr = h->opt.max_rings;
  bool regrown = false;
  int rc;
  if ((rc = ensure(h, h->d_scans, ns, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_point_base, ns + 1, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_ring16, np, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_tile_hist, std::max<uint64_t>(total_tiles, 1) * mr, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_rings, ns * mr, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_work, ns * mr, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_ring_featoff, ns * mr, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_ring_src, ns * mr, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_scan_flags, ns, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_gen_scan, ns, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_gen_tile_base, ns + 1, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_tile_owner, std::max<uint64_t>(total_tiles, 1), &regrown))) { return rc; }
  if (h->fast_enabled) {
    // a regular scan contributes at most min(n_points / FAST_MIN_RING, max_rings) rings to one list
    size_t fast_cap = 0;
    for (int s = 0; s < n_scans; s++) { fast_cap += std::min<size_t>(scans[s].n_points / FAST_MIN_RING, mr); }
    fast_cap = std::max<size_t>(fast_cap, 1);
    for (int c = 0; c < 2 * N_FAST_K; c++) {
      if ((rc = ensure(h, h->d_fast[c], fast_cap, &regrown))) { return rc; }
      if ((rc = ensure(h, h->d_rec[c], fast_cap * h->params.n_blocks, &regrown))) { return rc; }
    }
    for (int c = 0; c < N_FAST_K; c++) {
      if ((rc = ensure(h, h->d_bndx[c], fast_cap * FAST_BND, &regrown))) { return rc; }
    }
  }
  if ((rc = ensure(h, h->d_ring_path, ns * mr, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_idx, np, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_labels, np, &regrown))) { return rc; }
  if (h->opt.want_sorted_src && (rc = ensure(h, h->d_sorted_src, np, &regrown))) { return rc; }
  if (h->opt.want_curvature && (rc = ensure(h, h->d_curv, np, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_stage, np, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_edge, np, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_surface, np, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_counts, ns * 2, &regrown))) { return rc; }
  if ((rc = ensure(h, h->d_offsets, (ns + 1) * 2, &regrown))) { return rc; }
  if (input_bytes && (rc = ensure(h, h->d_input, input_bytes, &regrown))) { return rc; }
  if (regrown) { drop_graphs(h); }
  if (ns + 1 > h->h_scans_cap) {
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFreeHost(h->h_scans_buf); cudaFreeHost(h->h_point_base_buf);
    h->h_scans_buf = nullptr; h->h_point_base_buf = nullptr;
    h->h_scans = nullptr; h->h_point_base = nullptr;
    h->h_scans_cap = 0;
    const size_t want = ns + ns / 8 + 8;
    LFX_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_scans_buf), 2 * want * sizeof(ScanDesc)));
    LFX_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_point_base_buf), 2 * (want + 1) * sizeof(uint64_t)));
    h->h_scans_cap = want;
    for (auto & ev : h->desc_done) {
      if (!ev) { LFX_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
    }
  }
  // the slot written now was last read by the H2D copies of the batch before the previous one: wait for those
  // only, not for the stream (the previous batch keeps running while this one is laid out and enqueued)
  h->desc_slot ^= 1;
  LFX_CUDA(h, cudaEventSynchronize(h->desc_done[h->desc_slot]));
  h->h_scans = h->h_scans_buf + (size_t)h->desc_slot * h->h_scans_cap;
  h->h_point_base = h->h_point_base_buf + (size_t)h->desc_slot * (h->h_scans_cap + 1);

  // ---- descriptors + H2D of host-resident payloads (adjacent host buffers coalesce into one copy)
  uint64_t pb = 0;
  uint32_t tb = 0;
  size_t in_off = 0;
  const uint8_t * run_src = nullptr;
  size_t run_dst = 0, run_len = 0;
  auto flush = [&]() -> cudaError_t {
    if (!run_len) { return cudaSuccess; }
    cudaError_t e = cudaMemcpyAsync(h->d_input.p + run_dst, run_src, run_len, cudaMemcpyHostToDevice, h->stream);
    run_len = 0;
    return e;
  };
  for (int s = 0; s < n_scans; s++) {
    const lfx_cloud_view & v = scans[s];
    ScanDesc & d = h->h_scans[s];
    const size_t bytes = (size_t)v.n_points * v.point_step;
    if (v.memory == LFX_MEM_HOST) {
      d.data = h->d_input.p + in_off;
      if (bytes) {
        const uint8_t * src = static_cast<const uint8_t *>(v.data);
        if (run_len && run_src + run_len == src && run_dst + run_len == in_off) { run_len += bytes; }
        else { LFX_CUDA(h, flush()); run_src = src; run_dst = in_off; run_len = bytes; }
      }
      in_off += (bytes + 15) & ~(size_t)15;
    } else {
      d.data = static_cast<const uint8_t *>(v.data);
    }
    d.point_base = pb;
    d.n_points = v.n_points;
    d.point_step = v.point_step;
    d.off_x = v.off_x; d.off_y = v.off_y; d.off_z = v.off_z; d.off_ring = v.off_ring;
    d.ring_dt = v.ring_datatype;
    d.tile_base = tb;
    d.n_tiles = (v.n_points + h->tile - 1) / h->tile;
    d.ring16_lo = d.ring16_hi = 0;
    if (v.memory != LFX_MEM_HOST && h->have_conv && h->d_conv_out.p && v.point_step == 32 && v.off_ring == 20 && v.ring_datatype == LFX_RING_U16) {
      // a cloud the converter of this handle has just produced: its ring ids exist as a compact u16 array
      const uint8_t * p0 = static_cast<const uint8_t *>(v.data);
      const uint8_t * lo = h->d_conv_out.p, * hi = h->d_conv_out.p + h->conv_point_base.back() * 32;
      if (p0 >= lo && p0 + bytes <= hi && ((p0 - lo) & 31) == 0) {
        const uint64_t a16 = reinterpret_cast<uint64_t>(h->d_conv_ring16.p + ((p0 - lo) >> 5));
        d.ring16_lo = (uint32_t)a16; d.ring16_hi = (uint32_t)(a16 >> 32);
      }
    }
    d.vec_ok = (v.off_y == v.off_x + 4 && v.off_z == v.off_x + 8 && v.off_x % 16 == 0 && v.point_step % 16 == 0 &&
                v.off_x + 16 <= v.point_step && reinterpret_cast<uintptr_t>(d.data) % 16 == 0) ? 1u : 0u;
    h->h_point_base[s] = pb;
    pb += v.n_points;
    tb += d.n_tiles;
  }
  LFX_CUDA(h, flush());
  h->h_point_base[n_scans] = pb;
  if (n_scans > 0) {
    LFX_CUDA(h, cudaMemcpyAsync(h->d_scans.p, h->h_scans, sizeof(ScanDesc) * n_scans, cudaMemcpyHostToDevice, h->stream));
  }
  LFX_CUDA(h, cudaMemcpyAsync(h->d_point_base.p, h->h_point_base, sizeof(uint64_t) * (n_scans + 1), cudaMemcpyHostToDevice, h->stream));
  LFX_CUDA(h, cudaEventRecord(h->desc_done[h->desc_slot], h->stream));

  // ---- launch
  h->n_scans = n_scans;
  h->total_points = total_points;
  h->total_tiles = (uint32_t)total_tiles;
  if (n_scans > 0) {
    const bool timed_graph = h->timing && h->timing_in_graph;
    const bool use_graph = h->opt.use_graph >= 0 && (!h->timing || timed_graph);
    if (use_graph) {
      const GraphKey key{n_scans, (uint32_t)total_tiles, timed_graph};
      auto it = h->graphs.find(key);
      if (it == h->graphs.end()) {
        cudaGraph_t g = nullptr;
        LFX_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_pipeline(h, n_scans, (uint32_t)total_tiles, timed_graph, true);
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (rc) { if (g) { cudaGraphDestroy(g); } return rc; }
        if (e != cudaSuccess) { return fail(h, LFX_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
        cudaGraphExec_t ge = nullptr;
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { return fail(h, LFX_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
        if (h->graphs.size() >= 64) { drop_graphs(h); }
        it = h->graphs.emplace(key, ge).first;
      }
      LFX_CUDA(h, cudaGraphLaunch(it->second, h->stream));
      h->have_timing = timed_graph;
    } else {
      if ((rc = enqueue_pipeline(h, n_scans, (uint32_t)total_tiles, h->timing))) { return rc; }
      h->have_timing = h->timing;
    }
    h->launches += kernels_per_batch(h, use_graph && !timed_graph && h->cond_ok && n_scans >= COND_MIN_SCANS);
  } else {
    LFX_CUDA(h, cudaMemsetAsync(h->d_counters, 0, sizeof(uint32_t) * C_COUNT, h->stream));
    LFX_CUDA(h, cudaMemsetAsync(h->d_offsets.p, 0, sizeof(uint32_t) * 2, h->stream));
  }
  h->have_batch = true;

  if (out) {
    out->n_scans = n_scans;
    out->total_points = total_points;
    out->d_edge_xyz = reinterpret_cast<const float *>(h->d_edge.p);
    out->d_surface_xyz = reinterpret_cast<const float *>(h->d_surface.p);
    out->d_counts = h->d_counts.p;
    out->d_offsets = h->d_offsets.p;
    out->d_labels = h->d_labels.p;
    out->d_sorted_src = h->opt.want_sorted_src ? h->d_sorted_src.p : nullptr;
    out->d_curvature = h->opt.want_curvature ? h->d_curv.p : nullptr;
    out->d_rings = h->d_rings.p;
    out->d_point_base = h->d_point_base.p;
    out->max_rings = h->opt.max_rings;
  }
  return LFX_OK;
}

int lfx_synchronize(lfx_handle * h)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_batch_status(lfx_handle * h)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(uint32_t) * C_COUNT, cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->h_counters[C_ERR_FLAG]) {
    char buf[160];
    snprintf(buf, sizeof(buf), "capacity exceeded at scan %u, ring %u (raise lfx_options.max_ring_points / max_rings)",
             h->h_counters[C_ERR_SCAN], h->h_counters[C_ERR_RING]);
    return fail(h, (int)h->h_counters[C_ERR_FLAG], buf);
  }
  return LFX_OK;
}

int lfx_fetch_counts(lfx_handle * h, uint32_t * counts, uint32_t * offsets)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  if (counts && h->n_scans) { LFX_CUDA(h, cudaMemcpyAsync(counts, h->d_counts.p, sizeof(uint32_t) * 2 * h->n_scans, cudaMemcpyDeviceToHost, h->stream)); }
  if (offsets) { LFX_CUDA(h, cudaMemcpyAsync(offsets, h->d_offsets.p, sizeof(uint32_t) * 2 * (h->n_scans + 1), cudaMemcpyDeviceToHost, h->stream)); }
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_fetch_features(lfx_handle * h, float * edge_xyz, size_t edge_capacity_points, float * surface_xyz,
                       size_t surface_capacity_points)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  uint32_t tot[2];
  LFX_CUDA(h, cudaMemcpyAsync(tot, h->d_offsets.p + 2 * h->n_scans, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  if ((edge_xyz && tot[0] > edge_capacity_points) || (surface_xyz && tot[1] > surface_capacity_points)) {
    return fail(h, LFX_E_CAPACITY, "feature output buffer too small");
  }
  if (edge_xyz && tot[0]) { LFX_CUDA(h, cudaMemcpyAsync(edge_xyz, h->d_edge.p, sizeof(float4) * tot[0], cudaMemcpyDeviceToHost, h->stream)); }
  if (surface_xyz && tot[1]) { LFX_CUDA(h, cudaMemcpyAsync(surface_xyz, h->d_surface.p, sizeof(float4) * tot[1], cudaMemcpyDeviceToHost, h->stream)); }
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_fetch_points(lfx_handle * h, uint8_t * labels, uint32_t * sorted_src, double * curvature)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  if (sorted_src && !h->opt.want_sorted_src) { return fail(h, LFX_E_STATE, "handle was created without want_sorted_src"); }
  if (curvature && !h->opt.want_curvature) { return fail(h, LFX_E_STATE, "handle was created without want_curvature"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  const size_t n = h->total_points;
  if (n) {
    if (labels) { LFX_CUDA(h, cudaMemcpyAsync(labels, h->d_labels.p, n, cudaMemcpyDeviceToHost, h->stream)); }
    if (sorted_src) { LFX_CUDA(h, cudaMemcpyAsync(sorted_src, h->d_sorted_src.p, n * 4, cudaMemcpyDeviceToHost, h->stream)); }
    if (curvature) { LFX_CUDA(h, cudaMemcpyAsync(curvature, h->d_curv.p, n * 8, cudaMemcpyDeviceToHost, h->stream)); }
  }
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_fetch_rings(lfx_handle * h, lfx_ring_info * rings)
{
  if (!h || !rings) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  if (h->n_scans) {
    LFX_CUDA(h, cudaMemcpyAsync(rings, h->d_rings.p, sizeof(lfx_ring_info) * (size_t)h->n_scans * h->opt.max_rings, cudaMemcpyDeviceToHost, h->stream));
  }
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_extract_scan(lfx_handle * h, const lfx_cloud_view * scan, lfx_scan_output * out)
{
  if (!h || !scan || !out) { return LFX_E_BAD_PARAM; }
  int rc = lfx_extract_batch(h, scan, 1, nullptr);
  if (rc) { return rc; }
  const size_t n = std::max<size_t>(scan->n_points, 1);
  if (n > h->h_scan_cap) {
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFreeHost(h->h_edge); cudaFreeHost(h->h_surface); cudaFreeHost(h->h_labels); cudaFreeHost(h->h_sorted_src);
    h->h_edge = h->h_surface = nullptr; h->h_labels = nullptr; h->h_sorted_src = nullptr; h->h_scan_cap = 0;
    const size_t want = n + n / 8;
    LFX_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_edge), want * sizeof(float4)));
    LFX_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_surface), want * sizeof(float4)));
    LFX_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_labels), want));
    LFX_CUDA(h, cudaMallocHost(reinterpret_cast<void **>(&h->h_sorted_src), want * sizeof(uint32_t)));
    h->h_scan_cap = want;
  }
  uint32_t counts[2] = {0, 0};
  if ((rc = lfx_fetch_counts(h, counts, nullptr))) { return rc; }
  if ((rc = lfx_fetch_features(h, reinterpret_cast<float *>(h->h_edge), h->h_scan_cap, reinterpret_cast<float *>(h->h_surface), h->h_scan_cap))) { return rc; }
  if ((rc = lfx_fetch_points(h, h->h_labels, h->opt.want_sorted_src ? h->h_sorted_src : nullptr, nullptr))) { return rc; }
  if ((rc = lfx_batch_status(h))) { return rc; }
  out->edge_xyz = reinterpret_cast<const float *>(h->h_edge);
  out->surface_xyz = reinterpret_cast<const float *>(h->h_surface);
  out->n_edge = counts[0];
  out->n_surface = counts[1];
  out->labels = h->h_labels;
  out->sorted_src = h->opt.want_sorted_src ? h->h_sorted_src : nullptr;
  out->n_points = scan->n_points;
  return LFX_OK;
}

int lfx_label_to_color(uint8_t label, uint8_t * rgb)
{
  // LabelToColor, extraction/src/color_points.cpp:39-68
  static const uint8_t table[8][3] = {{255, 255, 255}, {255, 0, 0}, {255, 63, 0}, {255, 0, 0},
                                      {255, 63, 0}, {127, 127, 127}, {255, 0, 255}, {0, 255, 0}};
  if (label > 7 || !rgb) { return LFX_E_BAD_PARAM; }  // ThrowIfInvalidLabelDetected color_points.cpp:33-37
  memcpy(rgb, table[label], 3);
  return LFX_OK;
}

// ---------------------------------------------------------------- memory helpers

void * lfx_host_alloc(size_t bytes)
{
  void * p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { return nullptr; }
  return p;
}

void lfx_host_free(void * p) { if (p) { cudaFreeHost(p); } }

// NUMA node the handle's GPU hangs off (sysfs), or -1
static int gpu_numa_node(int device)
{
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
  std::string lower(bus);
  for (char & c : lower) { if (c >= 'A' && c <= 'F') { c = (char)(c - 'A' + 'a'); } }   // sysfs spells it in lower case
  const std::string path = "/sys/bus/pci/devices/" + lower + "/numa_node";
  int node = -1;
  if (FILE * f = fopen(path.c_str(), "r")) {
    if (fscanf(f, "%d", &node) != 1) { node = -1; }
    fclose(f);
  }
  if (node >= 0) { return node; }
  // Virtual machines often leave sysfs at -1 although the driver knows the node (the "NUMA Affinity" column of
  // `nvidia-smi topo -m`): ask NVML, loaded at run time so that the library does not link against it.
  void * so = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!so) { return -1; }
  using init_fn = int (*)();
  using by_bus_fn = int (*)(const char *, void **);
  using affinity_fn = int (*)(void *, unsigned int, unsigned long *, int);
  auto init = reinterpret_cast<init_fn>(dlsym(so, "nvmlInit_v2"));
  auto by_bus = reinterpret_cast<by_bus_fn>(dlsym(so, "nvmlDeviceGetHandleByPciBusId_v2"));
  auto affinity = reinterpret_cast<affinity_fn>(dlsym(so, "nvmlDeviceGetMemoryAffinity"));
  void * dev = nullptr;
  unsigned long set[16] = {0};
  if (init && by_bus && affinity && init() == 0 && by_bus(bus, &dev) == 0 && affinity(dev, 16u, set, 0 /* NVML_AFFINITY_SCOPE_NODE */) == 0) {
    for (int w = 0; w < 16 && node < 0; w++) {
      if (set[w]) { node = w * (int)(8 * sizeof(unsigned long)) + __builtin_ctzl(set[w]); }
    }
  }
  // (no nvmlShutdown / dlclose: other users of NVML in the process keep their reference)
  return node;
}

void * lfx_host_alloc_on(lfx_handle * h, size_t bytes, int * numa_node_out)
{
  // Pinned memory whose pages sit on the NUMA node of the handle's GPU: with one process per GPU on a two-socket host,
  // buffers that all end up on the node the processes happened to start on make every other GPU's DMA cross the
  // socket interconnect (r01s: 8 ranks reached 160 GB/s of host-to-device copies together, 51 GB/s alone).
  // set_mempolicy(MPOL_PREFERRED) around the allocation; the raw system call, so that libnuma is not a dependency.
  const int node = h ? gpu_numa_node(h->device) : -1;
  if (numa_node_out) { *numa_node_out = node; }
  bool bound = false;
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    bound = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024ul + 1) == 0;
  }
  void * p = nullptr;
  if (h) { cudaSetDevice(h->device); }
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { p = nullptr; cudaGetLastError(); }
  if (p) { memset(p, 0, bytes ? bytes : 1); }   // (pinning allocates the pages; this only makes sure of it while the policy holds)
  if (bound) { syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul); }
  if (numa_node_out && !bound) { *numa_node_out = -1; }
  return p;
}

int lfx_device_alloc(lfx_handle * h, size_t bytes, void ** out)
{
  if (!h || !out) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMalloc(out, bytes ? bytes : 1));
  return LFX_OK;
}

int lfx_device_free(lfx_handle * h, void * p)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  LFX_CUDA(h, cudaFree(p));
  return LFX_OK;
}

int lfx_memcpy_h2d(lfx_handle * h, void * dst, const void * src, size_t bytes)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_memcpy_d2h(lfx_handle * h, void * dst, const void * src, size_t bytes)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

// ---------------------------------------------------------------- instrumentation

uint64_t lfx_kernel_launch_count(const lfx_handle * h) { return h ? h->launches : 0; }

int lfx_set_stage_timing(lfx_handle * h, int enabled)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  h->timing = enabled != 0;
  h->timing_in_graph = enabled == 2;
  return LFX_OK;
}

int lfx_last_stage_ms(lfx_handle * h, float * ms)
{
  if (!h || !ms) { return LFX_E_BAD_PARAM; }
  if (!h->have_timing) { return fail(h, LFX_E_STATE, "stage timing was not enabled for the last batch"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaEventSynchronize(h->ev[LFX_N_STAGES]));
  for (int k = 0; k < LFX_N_STAGES; k++) { LFX_CUDA(h, cudaEventElapsedTime(&ms[k], h->ev[k], h->ev[k + 1])); }
  return LFX_OK;
}

int lfx_last_batch_stats(lfx_handle * h, lfx_batch_stats * out)
{
  if (!h || !out) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(uint32_t) * C_COUNT, cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  memset(out, 0, sizeof(*out));
  for (int c = 0; c < N_FAST_K; c++) { out->fast_rings[c] = h->h_counters[C_N_FAST0 + c]; }
  for (int c = 0; c < N_FAST_K; c++) { out->indexed_rings[c] = h->h_counters[C_N_FASTX0 + c]; }
  out->general_scans = h->h_counters[C_GEN_SCANS];
  out->general_rings = h->h_counters[C_N_WORK];
  return LFX_OK;
}

// ---------------------------------------------------------------- synthetic scans

int lfx_synth_named(const char * name, lfx_synth_spec * out) { return lfx_synth::named(name, out); }

int lfx_synth_scan_host(const lfx_synth_spec * spec, uint64_t frame, void * out, uint32_t * n_points_out)
{
  return lfx_synth::scan_host(spec, frame, out, n_points_out);
}

}  // extern "C"

namespace
{
__global__ void k_synth(const lfx_synth_spec spec, uint64_t first_frame, int n_scans, lfx_synth::Point32 * out)
{
  const size_t per_scan = (size_t)spec.n_rings * spec.n_cols;
  const size_t total = per_scan * n_scans;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const uint64_t scan = g / per_scan;
    const uint32_t k = (uint32_t)(g - scan * per_scan);
    const uint32_t col = k / spec.n_rings, ring = k - col * spec.n_rings;
    const uint64_t frame = first_frame + scan;
    const float az0 = lfx_synth::start_azimuth(spec, frame);
    lfx_synth::Point32 p;
    lfx_synth_spec s2 = spec;
    s2.dropout_prob = 0.0f;
    lfx_synth::make_point(s2, frame, az0, ring, col, &p);
    out[g] = p;
  }
}
}  // namespace

extern "C" int lfx_synth_batch_device(lfx_handle * h, const lfx_synth_spec * spec, uint64_t first_frame, int n_scans, void * d_out)
{
  if (!h || !spec || !d_out || n_scans < 0 || spec->n_rings <= 0 || spec->n_cols <= 0) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  if (n_scans == 0) { return LFX_OK; }
  k_synth<<<h->num_sms * 8, 256, 0, h->stream>>>(*spec, first_frame, n_scans, static_cast<lfx_synth::Point32 *>(d_out));
  LFX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return LFX_OK;
}

// ---------------------------------------------------------------- upstream converter

namespace
{

struct ConvField { std::string name; uint32_t offset; uint32_t dt; };

int conv_dt_size(uint32_t dt) { return dt <= 2 ? 1 : (dt <= 4 ? 2 : (dt <= 7 ? 4 : 8)); }

// Host half of PointTypeConverter.callback: convert.py:184-186 (append 'padding', stable sort by offset),
// create_point_format (convert.py:69-81: effective offsets never step backwards; the per-point format is
// max(end of last field, point_step) bytes long) and find_indices (convert.py:118-119). Returns LFX_CONVERT_*;
// `needs_six` reports that kept points could not be packed (decided once the kept count is known).
uint32_t conv_make_plan(const lfx_raw_cloud & rc, ConvCloud & cc, bool & needs_six, bool & ring_float)
{
  needs_six = false;
  ring_float = false;
  std::vector<ConvField> fs;
  for (uint32_t i = 0; i < rc.n_fields; i++) { fs.push_back({rc.fields[i].name ? rc.fields[i].name : "", rc.fields[i].offset, rc.fields[i].datatype}); }
  fs.push_back({"padding", 12u, 7u});
  std::stable_sort(fs.begin(), fs.end(), [](const ConvField & a, const ConvField & b) { return a.offset < b.offset; });
  std::vector<uint64_t> eff(fs.size());
  uint64_t index = 0;
  for (size_t i = 0; i < fs.size(); i++) {
    if (fs[i].dt < 1 || fs[i].dt > 8) { return LFX_CONVERT_E_DATATYPE; }
    if (index < fs[i].offset) { index = fs[i].offset; }
    eff[i] = index;
    index += (uint64_t)conv_dt_size(fs[i].dt);
  }
  const uint64_t n = rc.data_bytes / rc.point_step;
  if (n == 0) { return LFX_CONVERT_OK; }
  if (std::max<uint64_t>(index, rc.point_step) != rc.point_step) { return LFX_CONVERT_E_LAYOUT; }
  if (fs.size() < 3) { return LFX_CONVERT_E_FEW_FIELDS; }
  static const char * const keep_names[] = {"x", "y", "z", "padding", "intensity", "ring"};
  std::vector<size_t> retained;
  for (size_t i = 0; i < fs.size(); i++) {
    for (const char * k : keep_names) { if (fs[i].name == k) { retained.push_back(i); break; } }
  }
  const bool aligned_base = reinterpret_cast<uintptr_t>(rc.data) % 16 == 0 || rc.memory == LFX_MEM_HOST;  // host clouds are staged at aligned addresses
  auto put = [&](int slot, size_t i) {
    cc.off[slot] = (uint16_t)eff[i];
    cc.dt[slot] = (uint8_t)fs[i].dt;
    const uint64_t size = (uint64_t)conv_dt_size(fs[i].dt);
    cc.aligned[slot] = (aligned_base && eff[i] % size == 0 && rc.point_step % size == 0) ? 1 : 0;
  };
  for (int k = 0; k < 3; k++) { put(k, (size_t)k); }
  cc.packable = retained.size() == 6 ? 1 : 0;
  needs_six = !cc.packable;
  for (int k = 0; k < 6; k++) { put(3 + k, cc.packable ? retained[(size_t)k] : 0); }
  cc.fast = rc.is_bigendian ? 0 : 1;
  for (int k = 0; k < 8; k++) { if (cc.dt[k] != 7 || !cc.aligned[k]) { cc.fast = 0; } }
  if (!cc.aligned[8] || !(cc.dt[8] == 2 || cc.dt[8] == 4 || cc.dt[8] == 6)) { cc.fast = 0; }
  if (!cc.packable) { cc.fast = 0; }
  if (cc.packable && fs[retained[5]].dt >= 7) {   // 'H' wants an integer: the five float slots are still checked first
    ring_float = true;
    cc.dt[8] = 2; cc.off[8] = 0; cc.aligned[8] = 0;
  }
  return LFX_CONVERT_OK;
}

}  // namespace

extern "C" {

int lfx_convert_batch(lfx_handle * h, const lfx_raw_cloud * clouds, int n_clouds, lfx_convert_result * out)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (n_clouds < 0 || (n_clouds > 0 && !clouds)) { return fail(h, LFX_E_BAD_PARAM, "bad clouds argument"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  h->have_conv = false;
  h->conv_point_base.assign((size_t)n_clouds + 1, 0);
  h->conv_kept.assign((size_t)n_clouds, 0);
  h->conv_status.assign((size_t)n_clouds, LFX_CONVERT_OK);
  std::vector<ConvCloud> cc((size_t)n_clouds);
  std::vector<uint8_t> needs_six((size_t)n_clouds, 0), ring_float((size_t)n_clouds, 0), active((size_t)n_clouds, 0);
  std::vector<uint32_t> tile_cloud;
  bool all_fast = true;
  bool have_plan = false, plan_six = false, plan_rf = false, plan_aligned = false;
  lfx_raw_cloud plan_src{};
  ConvCloud plan_cc{};
  uint64_t points = 0, host_bytes = 0;
  uint32_t tiles = 0, max_step = 1;
  for (int c = 0; c < n_clouds; c++) {
    const lfx_raw_cloud & rc = clouds[c];
    ConvCloud & k = cc[(size_t)c];
    memset(&k, 0, sizeof(k));
    h->conv_point_base[(size_t)c] = points;
    k.tile_base = tiles;
    if (rc.n_fields > 0 && !rc.fields) { return fail(h, LFX_E_BAD_PARAM, "cloud without a field array"); }
    if (rc.point_step == 0 || rc.data_bytes % rc.point_step != 0) { h->conv_status[(size_t)c] = LFX_CONVERT_E_SIZE; continue; }
    if (rc.point_step > 65535u) { return fail(h, LFX_E_BAD_PARAM, "point_step > 65535 is outside the supported envelope"); }
    const uint64_t n = rc.data_bytes / rc.point_step;
    if (n > 0xFFFFFFFFull) { return fail(h, LFX_E_BAD_PARAM, "more than 2^32 - 1 points in one cloud"); }
    if (n > 0 && !rc.data) { return fail(h, LFX_E_BAD_PARAM, "cloud without data"); }
    bool six = false, rf = false;
    uint32_t st;
    // clouds of one driver share field table, point_step and byte order: resolve the plan once per batch
    const bool same_plan = have_plan && rc.fields == plan_src.fields && rc.n_fields == plan_src.n_fields && rc.point_step == plan_src.point_step &&
                           rc.is_bigendian == plan_src.is_bigendian && rc.memory == plan_src.memory && n > 0 &&
                           (rc.memory == LFX_MEM_HOST || reinterpret_cast<uintptr_t>(rc.data) % 16 == 0) == plan_aligned;
    if (same_plan) {
      k = plan_cc; six = plan_six; rf = plan_rf; st = LFX_CONVERT_OK;
      k.tile_base = tiles;
    } else {
      st = conv_make_plan(rc, k, six, rf);
      if (st == LFX_CONVERT_OK && n > 0) {
        have_plan = true; plan_src = rc; plan_cc = k; plan_six = six; plan_rf = rf;
        plan_aligned = rc.memory == LFX_MEM_HOST || reinterpret_cast<uintptr_t>(rc.data) % 16 == 0;
      }
    }
    h->conv_status[(size_t)c] = st;
    if (st != LFX_CONVERT_OK || n == 0) { continue; }
    needs_six[(size_t)c] = six; ring_float[(size_t)c] = rf; active[(size_t)c] = 1;
    all_fast = all_fast && k.fast;
    k.n_points = (uint32_t)n;
    k.point_step = rc.point_step;
    k.big = rc.is_bigendian ? 1 : 0;
    k.staged = ((uint64_t)rc.point_step * CONV_TILE <= (uint64_t)CONV_STAGE_MAX_BYTES && (rc.memory == LFX_MEM_HOST || reinterpret_cast<uintptr_t>(rc.data) % 16 == 0)) ? 1 : 0;
    if (k.staged) { max_step = std::max(max_step, rc.point_step); }
    if (rc.memory == LFX_MEM_HOST) { host_bytes = (host_bytes + 255) & ~255ull; host_bytes += rc.data_bytes; }
    points += n;
    const uint32_t nt = (uint32_t)((n + CONV_TILE - 1) / CONV_TILE);
    tile_cloud.insert(tile_cloud.end(), nt, (uint32_t)c);
    tiles += nt;
  }
  h->conv_point_base[(size_t)n_clouds] = points;
  int rc_all = LFX_OK;
  if (tiles > 0) {
    int rc;
    if ((rc = ensure(h, h->d_conv_raw, (size_t)host_bytes + 256, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_conv_out, (size_t)points * 32 + 32, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_conv_ring16, (size_t)points + 32, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_conv_clouds, (size_t)n_clouds, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_conv_state, (size_t)tiles, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_conv_tile_cloud, (size_t)tiles, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_conv_meta, (size_t)1 + 2 * (size_t)n_clouds, nullptr))) { return rc; }
    uint64_t hb = 0;
    for (int c = 0; c < n_clouds; c++) {
      if (!active[(size_t)c]) { continue; }
      ConvCloud & k = cc[(size_t)c];
      k.out = h->d_conv_out.p + h->conv_point_base[(size_t)c] * 32;
      if (clouds[c].memory == LFX_MEM_HOST) {
        hb = (hb + 255) & ~255ull;
        LFX_CUDA(h, cudaMemcpyAsync(h->d_conv_raw.p + hb, clouds[c].data, clouds[c].data_bytes, cudaMemcpyHostToDevice, h->stream));
        k.data = h->d_conv_raw.p + hb;
        hb += clouds[c].data_bytes;
      } else {
        k.data = static_cast<const uint8_t *>(clouds[c].data);
      }
    }
    LFX_CUDA(h, cudaMemcpyAsync(h->d_conv_clouds.p, cc.data(), sizeof(ConvCloud) * (size_t)n_clouds, cudaMemcpyHostToDevice, h->stream));
    LFX_CUDA(h, cudaMemcpyAsync(h->d_conv_tile_cloud.p, tile_cloud.data(), sizeof(uint32_t) * tiles, cudaMemcpyHostToDevice, h->stream));
    LFX_CUDA(h, cudaMemsetAsync(h->d_conv_state.p, 0, sizeof(unsigned long long) * tiles, h->stream));
    LFX_CUDA(h, cudaMemsetAsync(h->d_conv_meta.p, 0, sizeof(uint32_t) * (1 + 2 * (size_t)n_clouds), h->stream));
    ConvArgs a;
    a.clouds = h->d_conv_clouds.p;
    a.n_clouds = n_clouds;
    a.n_tiles = tiles;
    a.tile_cloud = h->d_conv_tile_cloud.p;
    a.tile_state = h->d_conv_state.p;
    a.ticket = h->d_conv_meta.p;
    a.kept = h->d_conv_meta.p + 1;
    a.flags = h->d_conv_meta.p + 1 + n_clouds;
    a.out_base = h->d_conv_out.p;
    a.ring16 = h->d_conv_ring16.p;
    if (!h->conv_smem_set) {
      LFX_CUDA(h, cudaFuncSetAttribute(k_convert<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_STAGE_MAX_BYTES));
      LFX_CUDA(h, cudaFuncSetAttribute(k_convert<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_STAGE_MAX_BYTES));
      for (auto & ev : h->conv_ev) { LFX_CUDA(h, cudaEventCreate(&ev)); }
      h->conv_smem_set = true;
    }
    LFX_CUDA(h, cudaEventRecord(h->conv_ev[0], h->stream));
    a.buf_bytes = ((uint32_t)CONV_TILE * max_step + 127u) & ~127u;
    if (all_fast) { k_convert<true><<<tiles, CONV_THREADS, (size_t)a.buf_bytes, h->stream>>>(a); }
    else { k_convert<false><<<tiles, CONV_THREADS, (size_t)a.buf_bytes, h->stream>>>(a); }
    LFX_CUDA(h, cudaGetLastError());
    LFX_CUDA(h, cudaEventRecord(h->conv_ev[1], h->stream));
    h->launches += 1;
    std::vector<uint32_t> meta((size_t)2 * n_clouds);
    LFX_CUDA(h, cudaMemcpyAsync(meta.data(), h->d_conv_meta.p + 1, sizeof(uint32_t) * 2 * (size_t)n_clouds, cudaMemcpyDeviceToHost, h->stream));
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int c = 0; c < n_clouds; c++) {
      if (!active[(size_t)c]) { continue; }
      const uint32_t kept = meta[(size_t)c], fl = meta[(size_t)n_clouds + c];
      uint32_t st = LFX_CONVERT_OK;
      if (kept > 0) {   // struct.pack only fails when there is something to pack
        if (needs_six[(size_t)c]) { st = LFX_CONVERT_E_FIELD_COUNT; }
        else if (fl & CONV_F_OVERFLOW) { st = LFX_CONVERT_E_OVERFLOW; }
        else if (ring_float[(size_t)c]) { st = LFX_CONVERT_E_RING_TYPE; }
        else if (fl & CONV_F_RING_RANGE) { st = LFX_CONVERT_E_RING_RANGE; }
      }
      h->conv_status[(size_t)c] = st;
      h->conv_kept[(size_t)c] = st == LFX_CONVERT_OK ? kept : 0;
    }
  }
  for (int c = 0; c < n_clouds; c++) {
    if (h->conv_status[(size_t)c] != LFX_CONVERT_OK && rc_all == LFX_OK) {
      rc_all = fail(h, LFX_E_CONVERT, "cloud " + std::to_string(c) + ": the reference converter raises here (LFX_CONVERT_* = " +
                                        std::to_string(h->conv_status[(size_t)c]) + ")");
    }
  }
  h->have_conv = true;
  if (out) {
    out->n_clouds = n_clouds;
    out->d_points = h->d_conv_out.p;
    out->point_base = h->conv_point_base.data();
    out->kept = h->conv_kept.data();
    out->status = h->conv_status.data();
  }
  return rc_all;
}

int lfx_last_convert_ms(lfx_handle * h, float * ms)
{
  if (!h || !ms) { return LFX_E_BAD_PARAM; }
  if (!h->have_conv || !h->conv_ev[0]) { return fail(h, LFX_E_STATE, "no batch has been converted"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaEventSynchronize(h->conv_ev[1]));
  LFX_CUDA(h, cudaEventElapsedTime(ms, h->conv_ev[0], h->conv_ev[1]));
  return LFX_OK;
}

int lfx_converted_view(lfx_handle * h, int cloud, lfx_cloud_view * out)
{
  if (!h || !out) { return LFX_E_BAD_PARAM; }
  if (!h->have_conv) { return fail(h, LFX_E_STATE, "no batch has been converted"); }
  if (cloud < 0 || (size_t)cloud >= h->conv_kept.size()) { return fail(h, LFX_E_BAD_PARAM, "cloud index out of range"); }
  if (h->conv_status[(size_t)cloud] != LFX_CONVERT_OK) { return fail(h, LFX_E_CONVERT, "this cloud was not converted"); }
  memset(out, 0, sizeof(*out));
  out->data = h->d_conv_out.p ? h->d_conv_out.p + h->conv_point_base[(size_t)cloud] * 32 : nullptr;
  out->n_points = h->conv_kept[(size_t)cloud];
  out->point_step = 32;                         // make_fields, convert.py:134-145
  out->off_x = 0; out->off_y = 4; out->off_z = 8; out->off_ring = 20;
  out->ring_datatype = LFX_RING_U16;
  out->has_ring = 1;
  out->is_dense = 1;                            // convert.py:210
  out->memory = LFX_MEM_DEVICE;
  return LFX_OK;
}

int lfx_converted_views(lfx_handle * h, lfx_cloud_view * out, int capacity, int * n_out)
{
  if (!h || !out || !n_out) { return LFX_E_BAD_PARAM; }
  if (!h->have_conv) { return fail(h, LFX_E_STATE, "no batch has been converted"); }
  int n = 0;
  for (size_t c = 0; c < h->conv_kept.size(); c++) {
    if (h->conv_status[c] != LFX_CONVERT_OK) { continue; }   // the reference's callback raises for this cloud: nothing is published
    if (n >= capacity) { return fail(h, LFX_E_CAPACITY, "more converted clouds than view slots"); }
    const int rc = lfx_converted_view(h, (int)c, &out[n]);
    if (rc != LFX_OK) { return rc; }
    n++;
  }
  *n_out = n;
  return LFX_OK;
}

int lfx_fetch_converted(lfx_handle * h, int cloud, void * dst, size_t capacity_bytes)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_conv) { return fail(h, LFX_E_STATE, "no batch has been converted"); }
  if (cloud < 0 || (size_t)cloud >= h->conv_kept.size()) { return fail(h, LFX_E_BAD_PARAM, "cloud index out of range"); }
  if (h->conv_status[(size_t)cloud] != LFX_CONVERT_OK) { return fail(h, LFX_E_CONVERT, "this cloud was not converted"); }
  const size_t bytes = (size_t)h->conv_kept[(size_t)cloud] * 32;
  if (bytes > capacity_bytes || (bytes > 0 && !dst)) { return fail(h, LFX_E_BAD_PARAM, "destination too small"); }
  if (bytes == 0) { return LFX_OK; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(dst, h->d_conv_out.p + h->conv_point_base[(size_t)cloud] * 32, bytes, cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- colored_scan + message layouts

extern "C" {

int lfx_color_batch(lfx_handle * h, lfx_colored_result * out)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  if (!h->opt.want_sorted_src) { return fail(h, LFX_E_STATE, "colored_scan needs lfx_options.want_sorted_src"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  const int ns = h->n_scans;
  h->have_colored = false;
  h->colored_counts.assign((size_t)ns, 0);
  h->colored_base.assign(h->h_point_base, h->h_point_base + ns + 1);
  if (ns > 0) {
    int rc;
    if ((rc = ensure(h, h->d_colored, (size_t)h->total_points * 2 + 2, nullptr))) { return rc; }
    if ((rc = ensure(h, h->d_colored_counts, (size_t)ns, nullptr))) { return rc; }
    ColorArgs a;
    a.scans = h->d_scans.p;
    a.rings = h->d_rings.p;
    a.labels = h->d_labels.p;
    a.sorted_src = h->d_sorted_src.p;
    a.out = h->d_colored.p;
    a.counts = h->d_colored_counts.p;
    a.max_rings = h->opt.max_rings;
    for (int s0 = 0; s0 < ns; s0 += 65535) {   // grid.y limit
      ColorArgs b = a;
      b.scans += s0; b.rings += (size_t)s0 * a.max_rings; b.counts += s0;
      k_color_scan<<<dim3((unsigned)a.max_rings, (unsigned)std::min(ns - s0, 65535)), COLOR_THREADS, 0, h->stream>>>(b);
      LFX_CUDA(h, cudaGetLastError());
      h->launches += 1;
    }
    LFX_CUDA(h, cudaMemcpyAsync(h->colored_counts.data(), h->d_colored_counts.p, sizeof(uint32_t) * (size_t)ns, cudaMemcpyDeviceToHost, h->stream));
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->have_colored = true;
  if (out) {
    out->n_scans = ns;
    out->d_points = reinterpret_cast<const uint8_t *>(h->d_colored.p);
    out->point_base = h->colored_base.data();
    out->counts = h->colored_counts.data();
  }
  return LFX_OK;
}

int lfx_fetch_colored(lfx_handle * h, int scan, void * dst, size_t capacity_bytes)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_colored) { return fail(h, LFX_E_STATE, "lfx_color_batch has not run for this batch"); }
  if (scan < 0 || (size_t)scan >= h->colored_counts.size()) { return fail(h, LFX_E_BAD_PARAM, "scan index out of range"); }
  const size_t bytes = (size_t)h->colored_counts[(size_t)scan] * 32;
  if (bytes > capacity_bytes || (bytes > 0 && !dst)) { return fail(h, LFX_E_BAD_PARAM, "destination too small"); }
  if (bytes == 0) { return LFX_OK; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(dst, reinterpret_cast<const uint8_t *>(h->d_colored.p) + h->colored_base[(size_t)scan] * 32, bytes, cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_topic_layout(int topic, lfx_point_field * fields, uint32_t * n_fields, uint32_t * point_step)
{
  if (!fields || !n_fields || !point_step) { return LFX_E_BAD_PARAM; }
  if (topic != LFX_TOPIC_SCAN_EDGE && topic != LFX_TOPIC_SCAN_SURFACE && topic != LFX_TOPIC_COLORED_SCAN) { return LFX_E_BAD_PARAM; }
  static const char * const names[4] = {"x", "y", "z", "rgb"};
  for (int k = 0; k < 3; k++) { fields[k].name = names[k]; fields[k].offset = 4u * (uint32_t)k; fields[k].datatype = 7; fields[k].count = 1; }
  *n_fields = 3;
  *point_step = 16;   // sizeof(pcl::PointXYZ)
  if (topic == LFX_TOPIC_COLORED_SCAN) {
    fields[3].name = names[3]; fields[3].offset = 16; fields[3].datatype = 7; fields[3].count = 1;
    *n_fields = 4;
    *point_step = 32; // sizeof(pcl::PointXYZRGB)
  }
  return LFX_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- mapping accumulate

namespace
{

// tf2::fromMsg(Pose, Isometry3d) = Translation3d(p) * Quaterniond(w, x, y, z), with Eigen's toRotationMatrix
void pose_matrix(const lfx_pose & p, double * m)
{
  const double x = p.orientation[0], y = p.orientation[1], z = p.orientation[2], w = p.orientation[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  m[0] = 1.0 - (tyy + tzz); m[1] = txy - twz; m[2] = txz + twy; m[3] = p.position[0];
  m[4] = txy + twz; m[5] = 1.0 - (txx + tzz); m[6] = tyz - twx; m[7] = p.position[1];
  m[8] = txz - twy; m[9] = tyz + twx; m[10] = 1.0 - (txx + tyy); m[11] = p.position[2];
}

// PoseDiffIsSufficientlySmall, map.hpp:50-60, on row-major 3x4 matrices: d = pose0^-1 * pose1, Quaterniond(d.rotation())
// by Eigen's matrix -> quaternion assignment, dt.norm() < tt && dq.vec().norm() < rt
bool pose_diff_small(const double * m0, const double * m1, double tt, double rt)
{
  double r0t[3][3], it[3], d[3][3], dt[3];
  for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) { r0t[a][b] = m0[4 * b + a]; } }
  for (int a = 0; a < 3; a++) { it[a] = -((r0t[a][0] * m0[3] + r0t[a][1] * m0[7]) + r0t[a][2] * m0[11]); }
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) { d[a][b] = (r0t[a][0] * m1[b] + r0t[a][1] * m1[4 + b]) + r0t[a][2] * m1[8 + b]; }
    dt[a] = ((r0t[a][0] * m1[3] + r0t[a][1] * m1[7]) + r0t[a][2] * m1[11]) + it[a];
  }
  double q[3] = {0.0, 0.0, 0.0};
  double t = d[0][0] + d[1][1] + d[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    t = 0.5 / t;
    q[0] = (d[2][1] - d[1][2]) * t; q[1] = (d[0][2] - d[2][0]) * t; q[2] = (d[1][0] - d[0][1]) * t;
  } else {
    int i = 0;
    if (d[1][1] > d[0][0]) { i = 1; }
    if (d[2][2] > d[i][i]) { i = 2; }
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(d[i][i] - d[j][j] - d[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[j] = (d[j][i] + d[i][j]) * t; q[k] = (d[k][i] + d[i][k]) * t;
  }
  const double dt_norm = std::sqrt((dt[0] * dt[0] + dt[1] * dt[1]) + dt[2] * dt[2]);
  const double dq_norm = std::sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
  return dt_norm < tt && dq_norm < rt;
}

// grow-only, contents preserved
int map_reserve(lfx_handle * h, uint64_t points)
{
  if (points <= h->d_map.cap) { return LFX_OK; }
  const size_t want = (size_t)points + (size_t)points / 2 + 4096;
  float4 * p = nullptr;
  LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(float4)));
  if (h->map_points) { LFX_CUDA(h, cudaMemcpyAsync(p, h->d_map.p, h->map_points * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream)); }
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->d_map.p) { LFX_CUDA(h, cudaFree(h->d_map.p)); }
  h->d_map.p = p;
  h->d_map.cap = want;
  return LFX_OK;
}

}  // namespace

extern "C" {

int lfx_pose_diff_is_small(const lfx_pose * pose0, const lfx_pose * pose1, double translation_threshold, double rotation_threshold)
{
  if (!pose0 || !pose1) { return -1; }
  double m0[12], m1[12];
  pose_matrix(*pose0, m0);
  pose_matrix(*pose1, m1);
  return pose_diff_small(m0, m1, translation_threshold, rotation_threshold) ? 1 : 0;
}

int lfx_map_gate(const lfx_pose * poses, const uint32_t * n_edge, int n, int * map_empty_io, lfx_pose * prev_io, uint8_t * selected)
{
  if (n < 0 || (n > 0 && (!poses || !n_edge)) || !map_empty_io || !prev_io) { return LFX_E_BAD_PARAM; }
  double prev[12] = {0}, m[12];
  bool empty = *map_empty_io != 0;
  if (!empty) { pose_matrix(*prev_io, prev); }
  for (int i = 0; i < n; i++) {
    if (selected) { selected[i] = 0; }
    if (n_edge[i] == 0) { continue; }
    pose_matrix(poses[i], m);
    if (!empty && pose_diff_small(prev, m, 1.0, 0.1)) { continue; }
    memcpy(prev, m, sizeof(m));
    *prev_io = poses[i];
    empty = false;
    if (selected) { selected[i] = 1; }
  }
  *map_empty_io = empty ? 1 : 0;
  return LFX_OK;
}

int lfx_map_set_state(lfx_handle * h, int map_empty, const lfx_pose * prev)
{
  if (!h || (!map_empty && !prev)) { return LFX_E_BAD_PARAM; }
  h->map_empty = map_empty != 0;
  if (!h->map_empty) { pose_matrix(*prev, h->map_prev); }
  return LFX_OK;
}

int lfx_map_add_batch(lfx_handle * h, const lfx_pose * poses, int n_poses, uint8_t * selected_out, uint64_t * map_points_out)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->have_batch) { return fail(h, LFX_E_STATE, "no batch has been extracted"); }
  if (n_poses != h->n_scans || (n_poses > 0 && !poses)) { return fail(h, LFX_E_BAD_PARAM, "one pose per scan of the last batch is required"); }
  LFX_CUDA(h, cudaSetDevice(h->device));
  const int ns = h->n_scans;
  std::vector<uint32_t> offsets((size_t)2 * (ns + 1), 0);
  if (ns > 0) {
    LFX_CUDA(h, cudaMemcpyAsync(offsets.data(), h->d_offsets.p, sizeof(uint32_t) * 2 * ((size_t)ns + 1), cudaMemcpyDeviceToHost, h->stream));
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  // MapBuilder::Callback, map.hpp:104-127, frame by frame (translation_threshold 1.0, rotation_threshold 0.1, :92-93)
  // The gate runs on copies of its state: the handle only advances once the frames it selected are in the map (a
  // failed reservation or launch must not leave later frames gated against a pose whose frame was never added).
  std::vector<MapFrame> frames;
  uint64_t dst = h->map_points;
  uint32_t longest = 0;
  bool gate_empty = h->map_empty;
  double gate_prev[12];
  memcpy(gate_prev, h->map_prev, sizeof(gate_prev));
  for (int s = 0; s < ns; s++) {
    if (selected_out) { selected_out[s] = 0; }
    const uint32_t n = offsets[2 * ((size_t)s + 1)] - offsets[2 * (size_t)s];
    if (n == 0) { continue; }                                                       // :117-120
    double m[12];
    pose_matrix(poses[s], m);
    if (!gate_empty && pose_diff_small(gate_prev, m, 1.0, 0.1)) { continue; }       // :122-128
    MapFrame f;
    memcpy(f.m, m, sizeof(m));
    f.dst = dst; f.src = offsets[2 * (size_t)s]; f.n = n;
    frames.push_back(f);
    dst += n;
    longest = std::max(longest, n);
    memcpy(gate_prev, m, sizeof(m));                                                // :131
    gate_empty = false;
    if (selected_out) { selected_out[s] = 1; }
  }
  if (!frames.empty()) {
    int rc;
    if ((rc = map_reserve(h, dst)) || (rc = ensure(h, h->d_map_frames, frames.size(), nullptr))) {
      if (selected_out) { memset(selected_out, 0, (size_t)ns); }   // nothing was added
      return rc;
    }
    LFX_CUDA(h, cudaMemcpyAsync(h->d_map_frames.p, frames.data(), sizeof(MapFrame) * frames.size(), cudaMemcpyHostToDevice, h->stream));
    const unsigned gy = (unsigned)std::min<uint32_t>(std::max<uint32_t>((longest + MAP_THREADS * 4 - 1) / (MAP_THREADS * 4), 1u), 64u);
    for (size_t f0 = 0; f0 < frames.size(); f0 += 1u << 30) {
      const unsigned gx = (unsigned)std::min<size_t>(frames.size() - f0, (size_t)1u << 30);
      k_map_transform_add<<<dim3(gx, gy), MAP_THREADS, 0, h->stream>>>(h->d_map_frames.p + f0, h->d_edge.p, h->d_map.p);
      LFX_CUDA(h, cudaGetLastError());
      h->launches += 1;
    }
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));   // `frames` is host memory of this call
    h->map_points = dst;
    h->map_empty = gate_empty;
    memcpy(h->map_prev, gate_prev, sizeof(gate_prev));
  }
  if (map_points_out) { *map_points_out = h->map_points; }
  return LFX_OK;
}

int lfx_map_size(lfx_handle * h, uint64_t * n_points_out)
{
  if (!h || !n_points_out) { return LFX_E_BAD_PARAM; }
  *n_points_out = h->map_points;
  return LFX_OK;
}

int lfx_map_fetch(lfx_handle * h, uint64_t first, uint64_t n_points, float * xyz)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (first + n_points > h->map_points || (n_points > 0 && !xyz)) { return fail(h, LFX_E_BAD_PARAM, "range outside the map"); }
  if (n_points == 0) { return LFX_OK; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LFX_CUDA(h, cudaMemcpyAsync(xyz, h->d_map.p + first, n_points * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

int lfx_map_clear(lfx_handle * h)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  h->map_points = 0;
  h->map_empty = true;
  return LFX_OK;
}

}  // extern "C"

// ====================================================================== multi-GPU driver

namespace
{

// NCCL is loaded at run time: the same libnccl.so.2 a host program (or torch) has already mapped is found first.
struct NcclApi
{
  void * lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char * (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi * nccl_api()
{
  static NcclApi api;
  if (api.lib || !api.err.empty()) { return &api; }
  for (const char * name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) { break; }
  }
  if (!api.lib) { api.err = "libnccl.so.2 could not be loaded (the multi-GPU driver needs NCCL)"; return &api; }
  auto sym = [&](const char * n) { void * p = dlsym(api.lib, n); if (!p && api.err.empty()) { api.err = std::string("NCCL lacks ") + n; } return p; };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return &api;
}

struct ShardWire   // what every rank tells every other rank at set-up
{
  cudaIpcMemHandle_t ipc;
  uint64_t pid;
  uint64_t ptr;       // the receive buffer's address in the owner's process (used when pid matches)
  int32_t device;
  int32_t pad;
};

}  // namespace

struct lfx_shard
{
  lfx_handle * h = nullptr;
  int rank = 0, world = 1;
  uint64_t n_frames = 0, first = 0, last = 0;
  uint32_t width = 0;             // rows of a block: the largest shard
  ncclComm_t comm = nullptr;
  int nranks = 0;
  bool p2p = false;
  uint32_t * d_recv = nullptr;    // SHARD_SLOTS receive slots, used round robin by epoch
  size_t slot_words = 0;
  uint32_t * peer_recv[lfxk::SHARD_MAX_WORLD] = {nullptr};   // peers' d_recv as mapped into this process
  bool peer_opened[lfxk::SHARD_MAX_WORLD] = {false};
  uint32_t * d_send = nullptr;    // NCCL mode: this rank's block, zero padded to width rows
  uint32_t * d_counts_all = nullptr;
  unsigned long long * d_offsets_all = nullptr;
  uint32_t * d_status = nullptr;
  uint32_t epoch = 0;             // exchanges started
  uint32_t finished = 0;          // exchanges whose scan has been enqueued
  std::string err;
};

namespace
{

int shard_fail(lfx_shard * s, int code, const std::string & msg) { s->err = msg; if (s->h) { s->h->err = msg; } return code; }

#define LFX_SHARD_CUDA(s, call)                                                                    \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) { return shard_fail(s, LFX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } \
  } while (0)
#define LFX_SHARD_NCCL(s, call)                                                                    \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) { return shard_fail(s, LFX_E_CUDA, std::string(#call) + ": " + nccl_api()->GetErrorString(r__)); } \
  } while (0)

// buffers + peer mapping, after s->comm exists. `wires`: null = gather them with NCCL (one process per rank)
int shard_setup(lfx_shard * s, const ShardWire * local_wires)
{
  NcclApi * nc = nccl_api();
  lfx_handle * h = s->h;
  LFX_SHARD_CUDA(s, cudaSetDevice(h->device));
  s->width = 0;
  for (int g = 0; g < s->world; g++) {
    const uint64_t a = ((uint64_t)g * s->n_frames) / (uint64_t)s->world, b = ((uint64_t)(g + 1) * s->n_frames) / (uint64_t)s->world;
    s->width = std::max<uint32_t>(s->width, (uint32_t)(b - a));
  }
  s->slot_words = lfxk::shard_slot_words(s->world, s->width);
  if (!local_wires) {   // (single-process groups allocate before they call this)
    LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->d_recv), sizeof(uint32_t) * lfxk::SHARD_SLOTS * s->slot_words));
    LFX_SHARD_CUDA(s, cudaMemset(s->d_recv, 0, sizeof(uint32_t) * lfxk::SHARD_SLOTS * s->slot_words));
  }
  LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->d_send), sizeof(uint32_t) * 2 * std::max<uint32_t>(s->width, 1)));
  LFX_SHARD_CUDA(s, cudaMemset(s->d_send, 0, sizeof(uint32_t) * 2 * std::max<uint32_t>(s->width, 1)));
  LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->d_counts_all), sizeof(uint32_t) * 2 * std::max<uint64_t>(s->n_frames, 1)));
  LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->d_offsets_all), sizeof(unsigned long long) * 2 * (s->n_frames + 1)));
  LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&s->d_status), sizeof(uint32_t)));
  LFX_SHARD_CUDA(s, cudaMemset(s->d_status, 0, sizeof(uint32_t)));
  LFX_SHARD_NCCL(s, nc->CommCount(s->comm, &s->nranks));
  // (a group inside one process always pushes: one host thread cannot issue a rank's ncclAllGather without the
  // other ranks' calls in the same NCCL group, which a per-rank entry point cannot provide)
  const char * mode = getenv("LFX_SHARD_EXCHANGE");
  const bool want_p2p = (local_wires || !(mode && strcmp(mode, "nccl") == 0)) && s->world <= lfxk::SHARD_MAX_WORLD;
  std::vector<ShardWire> wires((size_t)s->world);
  if (local_wires) {
    memcpy(wires.data(), local_wires, sizeof(ShardWire) * (size_t)s->world);
  } else {
    ShardWire mine;
    memset(&mine, 0, sizeof(mine));
    mine.pid = (uint64_t)getpid();
    mine.ptr = reinterpret_cast<uint64_t>(s->d_recv);
    mine.device = h->device;
    const bool ipc_ok = cudaIpcGetMemHandle(&mine.ipc, s->d_recv) == cudaSuccess;
    if (!ipc_ok) { cudaGetLastError(); mine.pid = 0; }   // pid 0: "cannot be mapped"
    ShardWire * d_w = nullptr;
    LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&d_w), sizeof(ShardWire) * ((size_t)s->world + 1)));
    LFX_SHARD_CUDA(s, cudaMemcpyAsync(d_w + s->world, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    LFX_SHARD_NCCL(s, nc->AllGather(d_w + s->world, d_w, sizeof(ShardWire), ncclChar, s->comm, h->stream));
    LFX_SHARD_CUDA(s, cudaMemcpyAsync(wires.data(), d_w, sizeof(ShardWire) * (size_t)s->world, cudaMemcpyDeviceToHost, h->stream));
    LFX_SHARD_CUDA(s, cudaStreamSynchronize(h->stream));
    cudaFree(d_w);
  }
  bool ok = want_p2p;
  for (int p = 0; p < s->world && ok; p++) {
    const ShardWire & w = wires[(size_t)p];
    if (p == s->rank) { s->peer_recv[p] = s->d_recv; continue; }
    if (w.pid == 0) { ok = false; break; }
    if (w.pid == (uint64_t)getpid()) {   // same process: plain peer access
      int can = 0;
      if (w.device != h->device) {
        if (cudaDeviceCanAccessPeer(&can, h->device, w.device) != cudaSuccess || !can) { cudaGetLastError(); ok = false; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(w.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = false; break; }
        cudaGetLastError();
      }
      s->peer_recv[p] = reinterpret_cast<uint32_t *>(w.ptr);
    } else {
      void * q = nullptr;
      if (cudaIpcOpenMemHandle(&q, w.ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      s->peer_recv[p] = static_cast<uint32_t *>(q);
      s->peer_opened[p] = true;
    }
  }
  // every rank must take the same path: agree through one more (tiny) all-gather, which is also the barrier after
  // which no rank's receive buffer is being zeroed any more
  if (!local_wires) {
    uint32_t * d_v = nullptr;
    LFX_SHARD_CUDA(s, cudaMalloc(reinterpret_cast<void **>(&d_v), sizeof(uint32_t) * ((size_t)s->world + 1)));
    const uint32_t v = ok ? 1u : 0u;
    LFX_SHARD_CUDA(s, cudaMemcpyAsync(d_v + s->world, &v, sizeof(v), cudaMemcpyHostToDevice, h->stream));
    LFX_SHARD_NCCL(s, nc->AllGather(d_v + s->world, d_v, 1, ncclUint32, s->comm, h->stream));
    std::vector<uint32_t> votes((size_t)s->world);
    LFX_SHARD_CUDA(s, cudaMemcpyAsync(votes.data(), d_v, sizeof(uint32_t) * (size_t)s->world, cudaMemcpyDeviceToHost, h->stream));
    LFX_SHARD_CUDA(s, cudaStreamSynchronize(h->stream));
    cudaFree(d_v);
    for (uint32_t x : votes) { ok = ok && x != 0; }
  }
  s->p2p = ok;
  return LFX_OK;
}

lfxk::ShardScanArgs shard_scan_args(lfx_shard * s)
{
  lfxk::ShardScanArgs a;
  a.slot = s->d_recv + (size_t)(s->epoch % lfxk::SHARD_SLOTS) * s->slot_words;
  a.counts_all = s->d_counts_all;
  a.offsets_all = s->d_offsets_all;
  a.status = s->d_status;
  a.n_frames = s->n_frames;
  a.world = s->world;
  a.width = s->width;
  a.epoch = s->p2p ? s->epoch : 0u;
  a.timeout_ns = 10ull * 1000 * 1000 * 1000;   // a peer that is 10 s late is gone: report instead of hanging the GPU
  return a;
}

int shard_enqueue_scan(lfx_shard * s)
{
  lfx_handle * h = s->h;
  lfxk::k_shard_scan<<<1, lfxk::SHARD_THREADS, 0, h->stream>>>(shard_scan_args(s));
  LFX_SHARD_CUDA(s, cudaGetLastError());
  h->launches += 1;
  s->finished = s->epoch;
  return LFX_OK;
}

}  // namespace

extern "C" {

int lfx_shard_range(uint64_t n_frames, int rank, int world, uint64_t * first, uint64_t * last)
{
  if (world <= 0 || rank < 0 || rank >= world || !first || !last) { return LFX_E_BAD_PARAM; }
  *first = ((uint64_t)rank * n_frames) / (uint64_t)world;
  *last = ((uint64_t)(rank + 1) * n_frames) / (uint64_t)world;
  return LFX_OK;
}

int lfx_shard_unique_id(void * id_out)
{
  if (!id_out) { return LFX_E_BAD_PARAM; }
  NcclApi * nc = nccl_api();
  if (!nc->err.empty()) { g_create_error = nc->err; return LFX_E_STATE; }
  static_assert(sizeof(ncclUniqueId) <= LFX_SHARD_ID_BYTES, "unique id fits the ABI's byte array");
  ncclUniqueId id;
  const ncclResult_t r = nc->GetUniqueId(&id);
  if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + nc->GetErrorString(r); return LFX_E_CUDA; }
  memset(id_out, 0, LFX_SHARD_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return LFX_OK;
}

int lfx_shard_create(lfx_handle * h, const void * unique_id, int rank, int world, uint64_t n_frames, lfx_shard ** out)
{
  if (!h || !out || world <= 0 || rank < 0 || rank >= world || (world > 1 && !unique_id)) { return LFX_E_BAD_PARAM; }
  NcclApi * nc = nccl_api();
  if (!nc->err.empty()) { return fail(h, LFX_E_STATE, nc->err); }
  lfx_shard * s = new lfx_shard();
  s->h = h; s->rank = rank; s->world = world; s->n_frames = n_frames;
  lfx_shard_range(n_frames, rank, world, &s->first, &s->last);
  if (cudaSetDevice(h->device) != cudaSuccess) { delete s; return fail(h, LFX_E_CUDA, "cudaSetDevice"); }
  ncclUniqueId id;
  if (world > 1) { memcpy(&id, unique_id, sizeof(id)); }
  else if (nc->GetUniqueId(&id) != ncclSuccess) { delete s; return fail(h, LFX_E_CUDA, "ncclGetUniqueId"); }
  const ncclResult_t r = nc->CommInitRank(&s->comm, world, id, rank);
  if (r != ncclSuccess) { const std::string m = std::string("ncclCommInitRank: ") + nc->GetErrorString(r); delete s; return fail(h, LFX_E_CUDA, m); }
  const int rc = shard_setup(s, nullptr);
  if (rc != LFX_OK) { lfx_shard_destroy(s); return rc; }
  *out = s;
  return LFX_OK;
}

int lfx_shard_create_local(lfx_handle ** handles, int world, uint64_t n_frames, lfx_shard ** out)
{
  if (!handles || !out || world <= 0 || world > lfxk::SHARD_MAX_WORLD) { return LFX_E_BAD_PARAM; }
  NcclApi * nc = nccl_api();
  if (!nc->err.empty()) { return fail(handles[0], LFX_E_STATE, nc->err); }
  std::vector<int> devs((size_t)world);
  for (int g = 0; g < world; g++) { if (!handles[g]) { return LFX_E_BAD_PARAM; } devs[(size_t)g] = handles[g]->device; }
  std::vector<ncclComm_t> comms((size_t)world, nullptr);
  const ncclResult_t r = nc->CommInitAll(comms.data(), world, devs.data());
  if (r != ncclSuccess) { return fail(handles[0], LFX_E_CUDA, std::string("ncclCommInitAll: ") + nc->GetErrorString(r)); }
  std::vector<ShardWire> wires((size_t)world);
  std::vector<lfx_shard *> made;
  int rc = LFX_OK;
  for (int g = 0; g < world; g++) {
    lfx_shard * s = new lfx_shard();
    made.push_back(s);
    s->h = handles[g]; s->rank = g; s->world = world; s->n_frames = n_frames; s->comm = comms[(size_t)g];
    lfx_shard_range(n_frames, g, world, &s->first, &s->last);
    uint32_t width = 0;
    for (int k = 0; k < world; k++) { uint64_t a, b; lfx_shard_range(n_frames, k, world, &a, &b); width = std::max<uint32_t>(width, (uint32_t)(b - a)); }
    const size_t words = lfxk::shard_slot_words(world, width);
    if (cudaSetDevice(s->h->device) != cudaSuccess || cudaMalloc(reinterpret_cast<void **>(&s->d_recv), sizeof(uint32_t) * lfxk::SHARD_SLOTS * words) != cudaSuccess ||
        cudaMemset(s->d_recv, 0, sizeof(uint32_t) * lfxk::SHARD_SLOTS * words) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      rc = fail(handles[0], LFX_E_CUDA, "receive buffer allocation failed");
      break;
    }
    memset(&wires[(size_t)g], 0, sizeof(ShardWire));
    wires[(size_t)g].pid = (uint64_t)getpid();
    wires[(size_t)g].ptr = reinterpret_cast<uint64_t>(s->d_recv);
    wires[(size_t)g].device = s->h->device;
  }
  bool all_p2p = true;
  for (int g = 0; g < world && rc == LFX_OK; g++) {
    rc = shard_setup(made[(size_t)g], wires.data());
    all_p2p = all_p2p && made[(size_t)g]->p2p;
  }
  if (rc == LFX_OK && !all_p2p) { rc = fail(handles[0], LFX_E_STATE, "the GPUs of a single-process group must be able to map each other's memory (peer access)"); }
  if (rc != LFX_OK) { for (lfx_shard * s : made) { lfx_shard_destroy(s); } return rc; }
  for (int g = 0; g < world; g++) { out[g] = made[(size_t)g]; }
  return LFX_OK;
}

int lfx_shard_exchange(lfx_shard * s)
{
  if (!s) { return LFX_E_BAD_PARAM; }
  lfx_handle * h = s->h;
  if (!h->have_batch) { return shard_fail(s, LFX_E_STATE, "no batch has been extracted"); }
  if ((uint64_t)h->n_scans != s->last - s->first) { return shard_fail(s, LFX_E_BAD_PARAM, "the last batch is not this rank's shard"); }
  LFX_SHARD_CUDA(s, cudaSetDevice(h->device));
  // the scan of the previous exchange goes first: after it this rank no longer reads the slot its peers write next
  const uint32_t n_local = (uint32_t)h->n_scans;
  if (s->p2p) {
    const bool scan_first = s->finished != s->epoch;
    const lfxk::ShardScanArgs sa = shard_scan_args(s);   // (of the previous epoch)
    s->epoch += 1;
    lfxk::ShardPushArgs pa;
    pa.counts = h->d_counts.p; pa.n_local = n_local; pa.rank = s->rank; pa.world = s->world; pa.width = s->width; pa.epoch = s->epoch;
    for (int p = 0; p < s->world; p++) { pa.peers.slot[p] = s->peer_recv[p] + (size_t)(s->epoch % lfxk::SHARD_SLOTS) * s->slot_words; }
    if (scan_first) { lfxk::k_shard_scan_push<<<1, lfxk::SHARD_THREADS, 0, h->stream>>>(sa, pa); s->finished = s->epoch - 1; }
    else { lfxk::k_shard_push<<<1, lfxk::SHARD_THREADS, 0, h->stream>>>(pa); }
    LFX_SHARD_CUDA(s, cudaGetLastError());
    h->launches += 1;
  } else {
    if (s->finished != s->epoch) { const int rc = shard_enqueue_scan(s); if (rc != LFX_OK) { return rc; } }
    s->epoch += 1;
    NcclApi * nc = nccl_api();
    LFX_SHARD_CUDA(s, cudaMemcpyAsync(s->d_send, h->d_counts.p, sizeof(uint32_t) * 2 * n_local, cudaMemcpyDeviceToDevice, h->stream));
    LFX_SHARD_NCCL(s, nc->AllGather(s->d_send, s->d_recv + (size_t)(s->epoch % lfxk::SHARD_SLOTS) * s->slot_words, (size_t)s->width * 2, ncclUint32, s->comm, h->stream));
  }
  return LFX_OK;
}

int lfx_shard_finish(lfx_shard * s, lfx_shard_result * out)
{
  if (!s || !out) { return LFX_E_BAD_PARAM; }
  if (s->epoch == 0) { return shard_fail(s, LFX_E_STATE, "no exchange has been started"); }
  LFX_SHARD_CUDA(s, cudaSetDevice(s->h->device));
  if (s->finished != s->epoch) { const int rc = shard_enqueue_scan(s); if (rc != LFX_OK) { return rc; } }
  out->n_frames = s->n_frames;
  out->first_frame = s->first;
  out->last_frame = s->last;
  out->d_counts_all = s->d_counts_all;
  out->d_offsets_all = reinterpret_cast<const uint64_t *>(s->d_offsets_all);
  return LFX_OK;
}

int lfx_shard_fetch(lfx_shard * s, uint32_t * counts_all, uint64_t * offsets_all)
{
  if (!s) { return LFX_E_BAD_PARAM; }
  lfx_shard_result r;
  const int rc = lfx_shard_finish(s, &r);
  if (rc != LFX_OK) { return rc; }
  lfx_handle * h = s->h;
  uint32_t status = 0;
  if (counts_all && s->n_frames) { LFX_SHARD_CUDA(s, cudaMemcpyAsync(counts_all, s->d_counts_all, sizeof(uint32_t) * 2 * s->n_frames, cudaMemcpyDeviceToHost, h->stream)); }
  if (offsets_all) { LFX_SHARD_CUDA(s, cudaMemcpyAsync(offsets_all, s->d_offsets_all, sizeof(uint64_t) * 2 * (s->n_frames + 1), cudaMemcpyDeviceToHost, h->stream)); }
  LFX_SHARD_CUDA(s, cudaMemcpyAsync(&status, s->d_status, sizeof(status), cudaMemcpyDeviceToHost, h->stream));
  LFX_SHARD_CUDA(s, cudaStreamSynchronize(h->stream));
  if (status != 0) { return shard_fail(s, LFX_E_STATE, "a peer rank did not publish its counts within 10 s"); }
  return LFX_OK;
}

int lfx_shard_info(const lfx_shard * s, int * uses_peer_stores, int * nccl_ranks)
{
  if (!s) { return LFX_E_BAD_PARAM; }
  if (uses_peer_stores) { *uses_peer_stores = s->p2p ? 1 : 0; }
  if (nccl_ranks) { *nccl_ranks = s->nranks; }
  return LFX_OK;
}

const char * lfx_shard_last_error(const lfx_shard * s) { return s ? s->err.c_str() : ""; }

void lfx_shard_destroy(lfx_shard * s)
{
  if (!s) { return; }
  if (s->h) { cudaSetDevice(s->h->device); if (s->h->stream) { cudaStreamSynchronize(s->h->stream); } }
  for (int p = 0; p < s->world && p < lfxk::SHARD_MAX_WORLD; p++) { if (s->peer_opened[p]) { cudaIpcCloseMemHandle(s->peer_recv[p]); } }
  cudaFree(s->d_recv); cudaFree(s->d_send); cudaFree(s->d_counts_all); cudaFree(s->d_offsets_all); cudaFree(s->d_status);
  if (s->comm && nccl_api()->CommDestroy) { nccl_api()->CommDestroy(s->comm); }
  delete s;
}

}  // extern "C"

// ====================================================================== localization residual build

struct LocMap
{
  double * x = nullptr, * y = nullptr, * z = nullptr;   // map order (what the neighbour indices address)
  uint64_t n = 0, cap = 0;
  // uniform grid over the map (k_loc_knn_grid): the same points cell by cell + their map indices
  double * gx = nullptr, * gy = nullptr, * gz = nullptr;
  uint32_t * gidx = nullptr, * cell_start = nullptr, * cell_count = nullptr;
  uint64_t cells_cap = 0;
  lfxk::LocGrid grid{};
  bool use_grid = false;
};

struct LocState
{
  LocMap map[2];
  DevBuf<float4> d_scan;
  DevBuf<uint32_t> d_nbr;
  DevBuf<double> d_d2, d_J, d_r;
};

namespace
{

LocState & loc_state(lfx_handle * h)
{
  if (!h->loc) { h->loc = new LocState(); }
  return *h->loc;
}

lfxk::LocPose loc_pose(const lfx_pose & p)
{
  lfxk::LocPose T;
  // Eigen::Quaterniond(w, x, y, z).toRotationMatrix()
  const double x = p.orientation[0], y = p.orientation[1], z = p.orientation[2], w = p.orientation[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T.r[0] = 1.0 - (tyy + tzz); T.r[1] = txy - twz; T.r[2] = txz + twy;
  T.r[3] = txy + twz; T.r[4] = 1.0 - (txx + tzz); T.r[5] = tyz - twx;
  T.r[6] = txz - twy; T.r[7] = tyz + twx; T.r[8] = 1.0 - (txx + tyy);
  T.t[0] = p.position[0]; T.t[1] = p.position[1]; T.t[2] = p.position[2];
  T.q[0] = x; T.q[1] = y; T.q[2] = z; T.q[3] = w;
  return T;
}

template<int K>
void launch_knn(lfx_handle * h, const LocMap & m, const float4 * scan, uint32_t n, const lfxk::LocPose & T, uint32_t * nbr, double * d2)
{
  const unsigned grid = (n + lfxk::LOC_KNN_WARPS - 1) / lfxk::LOC_KNN_WARPS;
  if (m.use_grid) { lfxk::k_loc_knn_grid<K><<<grid, lfxk::LOC_KNN_WARPS * 32, 0, h->stream>>>(m.grid, scan, n, T, nbr, d2); }
  else { lfxk::k_loc_knn<K><<<grid, lfxk::LOC_KNN_WARPS * 32, 0, h->stream>>>(m.x, m.y, m.z, (uint32_t)m.n, scan, n, T, nbr, d2); }
}

float loc_unkey(uint32_t k) { const uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k; float f; memcpy(&f, &b, 4); return f; }

// uniform grid over the map for k_loc_knn_grid (see lfx_loc.cuh); src: the map as float4 on the device
int loc_build_grid(lfx_handle * h, LocMap & m, const float4 * src)
{
  const char * ex = getenv("LFX_LOC_EXHAUSTIVE");   // diagnosis: answer by exhaustive search (k_loc_knn)
  m.use_grid = false;
  if (ex && atoi(ex) != 0) { return LFX_OK; }
  if (m.n > 0xFFFFFFF0ull) { return LFX_OK; }       // (refused later: 32-bit neighbour indices)
  uint32_t * d_keys = nullptr, h_keys[6];
  LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&d_keys), sizeof(h_keys)));
  cudaMemsetAsync(d_keys, 0xFF, 3 * sizeof(uint32_t), h->stream);
  cudaMemsetAsync(d_keys + 3, 0, 3 * sizeof(uint32_t), h->stream);
  lfxk::k_loc_bbox<<<h->num_sms * 4, 256, 0, h->stream>>>(src, m.n, d_keys);
  cudaMemcpyAsync(h_keys, d_keys, sizeof(h_keys), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d_keys);
  LFX_CUDA(h, e);
  lfxk::LocGrid g{};
  double ext[3] = {0, 0, 0};
  double lo[3] = {0, 0, 0};
  for (int a = 0; a < 3; a++) {
    if (h_keys[a] <= h_keys[3 + a]) { lo[a] = (double)loc_unkey(h_keys[a]); ext[a] = (double)loc_unkey(h_keys[3 + a]) - lo[a]; }
  }
  const char * cs = getenv("LFX_LOC_CELL");
  double c = cs && atof(cs) > 0 ? atof(cs) : 1.0;   // metres: LOAM maps are voxel-filtered to a few points per cell of this size
  uint64_t dims[3];
  for (;;) {
    uint64_t cells = 1;
    bool fits = true;
    for (int a = 0; a < 3; a++) {
      const double d = std::floor(ext[a] / c) + 1.0;
      if (!(d < 4.0e9)) { fits = false; break; }
      dims[a] = (uint64_t)d;
      cells *= dims[a];
      if (cells > lfxk::LOC_MAX_CELLS) { fits = false; break; }
    }
    if (fits) { break; }
    c *= 1.2599210498948732;   // 2^(1/3): half as many cells
  }
  g.x0 = lo[0]; g.y0 = lo[1]; g.z0 = lo[2];
  g.c = c; g.inv_c = 1.0 / c;
  g.nx = (int)dims[0]; g.ny = (int)dims[1]; g.nz = (int)dims[2];
  const uint64_t n_cells = dims[0] * dims[1] * dims[2];
  if (n_cells + 1 > m.cells_cap) {
    cudaFree(m.cell_start); cudaFree(m.cell_count);
    m.cell_start = m.cell_count = nullptr; m.cells_cap = 0;
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.cell_start), sizeof(uint32_t) * (n_cells + 1)));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.cell_count), sizeof(uint32_t) * (n_cells + 1)));
    m.cells_cap = n_cells + 1;
  }
  g.cell_start = m.cell_start; g.gx = m.gx; g.gy = m.gy; g.gz = m.gz; g.gidx = m.gidx;
  LFX_CUDA(h, cudaMemsetAsync(m.cell_count, 0, sizeof(uint32_t) * n_cells, h->stream));
  lfxk::k_loc_cell_count<<<h->num_sms * 4, 256, 0, h->stream>>>(m.x, m.y, m.z, m.n, g, m.cell_count);
  lfxk::k_loc_cell_scan<<<1, 1024, 0, h->stream>>>(m.cell_count, m.cell_start, (uint32_t)n_cells);
  lfxk::k_loc_cell_fill<<<h->num_sms * 4, 256, 0, h->stream>>>(m.x, m.y, m.z, m.n, g, m.cell_count, m.gx, m.gy, m.gz, m.gidx);
  LFX_CUDA(h, cudaGetLastError());
  h->launches += 4;
  m.grid = g;
  m.use_grid = true;
  return LFX_OK;
}

int loc_run(lfx_handle * h, int kind, const float * scan_xyz4, uint32_t n, int memory, const lfx_pose * pose, int k, double * J_out,
            double * r_out, uint32_t * nbr_out)
{
  if (!h || !pose || (n > 0 && !scan_xyz4) || (kind != LFX_LOC_EDGE && kind != LFX_LOC_SURFACE)) { return LFX_E_BAD_PARAM; }
  if (k < 1 || k > lfxk::LOC_MAX_K) { return fail(h, LFX_E_BAD_PARAM, "n_neighbors must be in 1..16"); }
  if (kind == LFX_LOC_SURFACE && k < 3) { return fail(h, LFX_E_BAD_PARAM, "a plane needs at least three neighbours"); }
  LocState & st = loc_state(h);
  const LocMap & m = st.map[kind];
  if (m.n < (uint64_t)k) { return fail(h, LFX_E_STATE, "the map holds fewer points than n_neighbors (lfx_loc_set_map)"); }
  if (m.n > 0xFFFFFFF0ull) { return fail(h, LFX_E_CAPACITY, "map too large for 32-bit neighbour indices"); }
  if (n == 0) { return LFX_OK; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  int rc;
  const float4 * d_scan = reinterpret_cast<const float4 *>(scan_xyz4);
  if (memory == LFX_MEM_HOST) {
    if ((rc = ensure(h, st.d_scan, n, nullptr))) { return rc; }
    LFX_CUDA(h, cudaMemcpyAsync(st.d_scan.p, scan_xyz4, sizeof(float4) * n, cudaMemcpyHostToDevice, h->stream));
    d_scan = st.d_scan.p;
  }
  const int jw = kind == LFX_LOC_EDGE ? 21 : 7, rw = kind == LFX_LOC_EDGE ? 3 : 1;
  if ((rc = ensure(h, st.d_nbr, (size_t)n * k, nullptr)) || (rc = ensure(h, st.d_d2, (size_t)n * k, nullptr)) ||
      (rc = ensure(h, st.d_J, (size_t)n * jw, nullptr)) || (rc = ensure(h, st.d_r, (size_t)n * rw, nullptr))) { return rc; }
  const lfxk::LocPose T = loc_pose(*pose);
  switch (k) {   // the per-lane candidate lists live in registers: k is a template parameter
#define LFX_KNN_CASE(KK) case KK: launch_knn<KK>(h, m, d_scan, n, T, st.d_nbr.p, st.d_d2.p); break;
    LFX_KNN_CASE(1) LFX_KNN_CASE(2) LFX_KNN_CASE(3) LFX_KNN_CASE(4) LFX_KNN_CASE(5) LFX_KNN_CASE(6) LFX_KNN_CASE(7) LFX_KNN_CASE(8)
    LFX_KNN_CASE(9) LFX_KNN_CASE(10) LFX_KNN_CASE(11) LFX_KNN_CASE(12) LFX_KNN_CASE(13) LFX_KNN_CASE(14) LFX_KNN_CASE(15) LFX_KNN_CASE(16)
#undef LFX_KNN_CASE
  }
  LFX_CUDA(h, cudaGetLastError());
  lfxk::LocArgs a;
  a.mx = m.x; a.my = m.y; a.mz = m.z; a.scan = d_scan; a.nbr = st.d_nbr.p; a.n = n; a.k = k; a.T = T; a.J = st.d_J.p; a.r = st.d_r.p;
  if (kind == LFX_LOC_EDGE) { lfxk::k_loc_edge<<<(n + 127) / 128, 128, 0, h->stream>>>(a); }
  else { lfxk::k_loc_surface<<<(n + 127) / 128, 128, 0, h->stream>>>(a); }
  LFX_CUDA(h, cudaGetLastError());
  h->launches += 2;
  if (J_out) { LFX_CUDA(h, cudaMemcpyAsync(J_out, st.d_J.p, sizeof(double) * (size_t)n * jw, cudaMemcpyDeviceToHost, h->stream)); }
  if (r_out) { LFX_CUDA(h, cudaMemcpyAsync(r_out, st.d_r.p, sizeof(double) * (size_t)n * rw, cudaMemcpyDeviceToHost, h->stream)); }
  if (nbr_out) { LFX_CUDA(h, cudaMemcpyAsync(nbr_out, st.d_nbr.p, sizeof(uint32_t) * (size_t)n * k, cudaMemcpyDeviceToHost, h->stream)); }
  LFX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LFX_OK;
}

}  // namespace

extern "C" {

int lfx_loc_set_map(lfx_handle * h, int kind, const float * xyz4, uint64_t n, int memory)
{
  if (!h || (n > 0 && !xyz4) || (kind != LFX_LOC_EDGE && kind != LFX_LOC_SURFACE)) { return LFX_E_BAD_PARAM; }
  LFX_CUDA(h, cudaSetDevice(h->device));
  LocState & st = loc_state(h);
  LocMap & m = st.map[kind];
  if (n > m.cap) {
    LFX_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(m.x); cudaFree(m.y); cudaFree(m.z); cudaFree(m.gx); cudaFree(m.gy); cudaFree(m.gz); cudaFree(m.gidx);
    m.x = m.y = m.z = m.gx = m.gy = m.gz = nullptr; m.gidx = nullptr; m.cap = 0;
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.x), sizeof(double) * n));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.y), sizeof(double) * n));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.z), sizeof(double) * n));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.gx), sizeof(double) * n));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.gy), sizeof(double) * n));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.gz), sizeof(double) * n));
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&m.gidx), sizeof(uint32_t) * n));
    m.cap = n;
  }
  m.n = n;
  m.use_grid = false;
  if (n == 0) { return LFX_OK; }
  const float4 * src = reinterpret_cast<const float4 *>(xyz4);
  float4 * tmp = nullptr;
  if (memory == LFX_MEM_HOST) {
    LFX_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&tmp), sizeof(float4) * n));
    LFX_CUDA(h, cudaMemcpyAsync(tmp, xyz4, sizeof(float4) * n, cudaMemcpyHostToDevice, h->stream));
    src = tmp;
  }
  lfxk::k_loc_soa<<<h->num_sms * 4, 256, 0, h->stream>>>(src, n, m.x, m.y, m.z);
  LFX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  const int grc = loc_build_grid(h, m, src);
  cudaError_t se = cudaStreamSynchronize(h->stream);
  if (tmp) { cudaFree(tmp); }
  if (grc != LFX_OK) { return grc; }
  LFX_CUDA(h, se);
  return LFX_OK;
}

int lfx_loc_edge(lfx_handle * h, const float * scan_xyz4, uint32_t n, int memory, const lfx_pose * point_to_map, int n_neighbors,
                 double * jacobians, double * residuals, uint32_t * neighbors)
{
  return loc_run(h, LFX_LOC_EDGE, scan_xyz4, n, memory, point_to_map, n_neighbors, jacobians, residuals, neighbors);
}

int lfx_loc_surface(lfx_handle * h, const float * scan_xyz4, uint32_t n, int memory, const lfx_pose * point_to_map, int n_neighbors,
                    double * jacobians, double * residuals, uint32_t * neighbors)
{
  return loc_run(h, LFX_LOC_SURFACE, scan_xyz4, n, memory, point_to_map, n_neighbors, jacobians, residuals, neighbors);
}

int lfx_loc_release(lfx_handle * h)
{
  if (!h) { return LFX_E_BAD_PARAM; }
  if (!h->loc) { return LFX_OK; }
  cudaSetDevice(h->device);
  if (h->stream) { cudaStreamSynchronize(h->stream); }
  LocState & st = *h->loc;
  for (LocMap & m : st.map) { cudaFree(m.x); cudaFree(m.y); cudaFree(m.z); cudaFree(m.gx); cudaFree(m.gy); cudaFree(m.gz); cudaFree(m.gidx); cudaFree(m.cell_start); cudaFree(m.cell_count); }
  cudaFree(st.d_scan.p); cudaFree(st.d_nbr.p); cudaFree(st.d_d2.p); cudaFree(st.d_J.p); cudaFree(st.d_r.p);
  delete h->loc;
  h->loc = nullptr;
  return LFX_OK;
}

}  // extern "C"
