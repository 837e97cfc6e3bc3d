// tma_gather_probe.cu — can tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) do the ring-sector gather?
// Geometry of the bench workload (like gather_probe.cu): scans of R = 128 rings x W = 2048 firings, 32-byte points in
// firing order, so ring r of a scan is a 3-D tensor slice {8 words, ring r, column c}: rows of 32 B that lie R * 32 B
// = 4 KB apart. One warp per (ring, sector) item lands 32 * K = 352 columns of its ring in ONE landing buffer
// (32 B per position), consumes x and the ring word of each of its K positions per lane, and only then lets lane 0
// issue the next item's boxes (single buffered, like the planned kernel: the copy has the whole compute phase of the
// item to land). `delay` cycles of spinning stand in for the compute phase.
//   mode 0: boxes of 32 columns (11 per item)     mode 1: boxes of 176 columns (2 per item)
//   mode 2: boxes of 176 columns, inner 16 B only (x,y,z,pad)   [half the landing bytes]
//   mode 3: like 1, tensor maps built ON THE DEVICE with tensormap.replace from one host-encoded template
// Every landed point is checked against its expected first word (= its global point index).
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DPROBE_NW=16 -o tools/probes/tma_gather_probe tools/probes/tma_gather_probe.cu
//   run:   tools/probes/tma_gather_probe [scans=1250] [delay_cycles=0]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef PROBE_NW
#define PROBE_NW 16
#endif
constexpr int R = 128, W = 2048, K = 11, NW = PROBE_NW, B = 6, POS = 32 * K;

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(void * dst, const CUtensorMap * map, int c0, int c1, int c2, uint64_t * bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// one thread block per scan: template -> smem, replace address / dims / strides, publish to global memory
__global__ void k_build_maps(const CUtensorMap * tmpl, CUtensorMap * maps, const uint8_t * in, int n_scans)
{
  __shared__ __align__(128) CUtensorMap sm;
  const int scan = blockIdx.x;
  if (threadIdx.x < 32) {
    reinterpret_cast<uint32_t *>(&sm)[threadIdx.x] = reinterpret_cast<const uint32_t *>(tmpl)[threadIdx.x];
    __syncwarp();
    if (threadIdx.x == 0) {
      const uint64_t sa = (uint64_t)smem_u32(&sm);
      const uint8_t * base = in + (size_t)scan * W * R * 32;
      asm volatile("tensormap.replace.tile.global_address.shared::cta.b1024.b64 [%0], %1;" :: "l"(sa), "l"(base) : "memory");
      asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 1, %1;" :: "l"(sa), "r"(R) : "memory");
      asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 2, %1;" :: "l"(sa), "r"(W) : "memory");
      asm volatile("tensormap.replace.tile.global_stride.shared::cta.b1024.b64 [%0], 0, %1;" :: "l"(sa), "l"((uint64_t)32) : "memory");
      asm volatile("tensormap.replace.tile.global_stride.shared::cta.b1024.b64 [%0], 1, %1;" :: "l"(sa), "l"((uint64_t)R * 32) : "memory");
    }
    __syncwarp();
    asm volatile("tensormap.cp_fenceproxy.global.shared::cta.tensormap::generic.release.gpu.sync.aligned [%0], [%1], 128;"
                 :: "l"(maps + scan), "r"(smem_u32(&sm)) : "memory");
  }
}

template<int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k_gather(const CUtensorMap * maps, int n_scans, int delay, int skew, unsigned long long * out)
{
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int SLOT = MODE == 2 ? 16 : 32;
  constexpr int CB = MODE == 0 ? 32 : 176;
  constexpr int PER_WARP = POS * SLOT + 128;   // one landing buffer + the barrier
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t * buf = smem + (size_t)warp * PER_WARP;
  uint64_t * bar = reinterpret_cast<uint64_t *>(buf + POS * SLOT);
  if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint32_t n_rings = (uint32_t)n_scans * R;
  const uint32_t n_units = ((n_rings + NW - 1) / NW) * B;
  unsigned long long bad = 0;
  uint32_t phase = 0;
  auto coords = [&](uint32_t unit, uint32_t & scan, uint32_t & r, uint32_t & c0) -> bool {
    const uint32_t chunk = unit / B, j = unit % B;
    const uint32_t e = chunk * NW + warp;
    scan = e / R; r = e % R;
    c0 = j * ((W - POS) / (B - 1));
    return unit < n_units && e < n_rings;
  };
  uint32_t last_scan = 0xFFFFFFFFu;
  auto issue = [&](uint32_t unit) {
    uint32_t scan, r, c0;
    if (!coords(unit, scan, r, c0)) { return; }
    if (lane == 0) {
      const CUtensorMap * map = maps + scan;
      if (MODE == 3 && scan != last_scan) {
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(map) : "memory");
        last_scan = scan;
      }
      mbar_expect(bar, POS * SLOT);
#pragma unroll
      for (int b = 0; b < POS / CB; b++) { tma_load_3d(buf + b * CB * SLOT, map, 0, (int)r, (int)c0 + b * CB, bar); }
    }
  };
  issue(blockIdx.x);
  for (uint32_t t = 0; blockIdx.x + t * gridDim.x < n_units; t++) {
    const uint32_t unit = blockIdx.x + t * gridDim.x;
    uint32_t scan, r, c0;
    const bool ok = coords(unit, scan, r, c0);
    if (ok) { mbar_wait(bar, phase); phase ^= 1; }
    uint32_t v[K], w[K];
    if (ok) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        v[k] = *reinterpret_cast<const uint32_t *>(buf + (lane * K + k) * SLOT);
        w[k] = MODE == 2 ? 0u : *reinterpret_cast<const uint32_t *>(buf + (lane * K + k) * SLOT + 20);
      }
      uint32_t acc = 0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint32_t expect = (scan * W + c0 + lane * K + k) * R + r;
        bad += v[k] != expect;
        acc |= w[k];
      }
      bad += acc != 0x01010101u && MODE != 2;
    }
    // the landing buffer is free once every lane holds its values (the vote consumes them)
    __syncwarp();
    issue(unit + gridDim.x);
    if (delay > 0) {
      // skew: warp w of a CTA spends (w * skew) / 16 percent longer, so that neighbouring rings drift apart
      const long long t0 = clock64(), d = delay + (long long)delay * warp * skew / 1600;
      while (clock64() - t0 < d) { }
    }
  }
  if (bad) { atomicAdd(out, bad); }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode()
{
  void * fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    exit(1);
  }
  return (EncodeFn)fn;
}

static CUtensorMap encode(EncodeFn enc, const uint8_t * base, int inner_words, int cb, int rings, int cols)
{
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)inner_words, (cuuint64_t)rings, (cuuint64_t)cols};
  cuuint64_t strides[2] = {32, (cuuint64_t)rings * 32};
  cuuint32_t box[3] = {(cuuint32_t)inner_words, 1, (cuuint32_t)cb};
  cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", (int)r); exit(1); }
  return m;
}

__global__ void k_fill(uint32_t * in, size_t n_points)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_points; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t * p = in + i * 8;
    p[0] = (uint32_t)i; p[1] = 0; p[2] = 0; p[3] = 0; p[4] = 0; p[5] = 0x01010101u; p[6] = 0; p[7] = 0;
  }
}

template<int MODE>
float run(const CUtensorMap * d_maps, int n_scans, int delay, int skew, unsigned long long * d_out, int sms, unsigned long long & bad)
{
  constexpr int SLOT = MODE == 2 ? 16 : 32;
  const size_t smem = (size_t)NW * (POS * SLOT + 128);
  cudaFuncSetAttribute(k_gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(d_out, 0, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) { k_gather<MODE><<<sms, NW * 32, smem>>>(d_maps, n_scans, delay, skew, d_out); }
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) { k_gather<MODE><<<sms, NW * 32, smem>>>(d_maps, n_scans, delay, skew, d_out); }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(err)); return -1.f; }
  cudaMemcpy(&bad, d_out, 8, cudaMemcpyDeviceToHost);
  return ms / 5;
}

int main(int argc, char ** argv)
{
  const int n_scans = argc > 1 ? atoi(argv[1]) : 1250;
  const int delay = argc > 2 ? atoi(argv[2]) : 0;
  const int skew = argc > 3 ? atoi(argv[3]) : 0;      // percent by which the last warp of a CTA is slower than the first
  const int gran = argc > 4 ? atoi(argv[4]) : 0;      // cudaLimitMaxL2FetchGranularity (0: leave the default)
  {
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity default %zu", g);
    if (gran) { const cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf(" -> %zu (%s)", g, cudaGetErrorString(e)); }
    printf("\n");
  }
  const size_t n_points = (size_t)n_scans * R * W, bytes = n_points * 32;
  uint8_t * d_in; unsigned long long * d_out;
  cudaMalloc(&d_in, bytes); cudaMalloc(&d_out, 8);
  k_fill<<<148 * 8, 256>>>(reinterpret_cast<uint32_t *>(d_in), n_points);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  EncodeFn enc = get_encode();
  std::vector<CUtensorMap> h32(n_scans), h176(n_scans), h176x(n_scans);
  for (int s = 0; s < n_scans; s++) {
    const uint8_t * base = d_in + (size_t)s * W * R * 32;
    h32[s] = encode(enc, base, 8, 32, R, W);
    h176[s] = encode(enc, base, 8, 176, R, W);
    h176x[s] = encode(enc, base, 4, 176, R, W);
  }
  CUtensorMap * d32, * d176, * d176x, * ddev, * dtmpl;
  const size_t mb = sizeof(CUtensorMap) * n_scans;
  cudaMalloc(&d32, mb); cudaMalloc(&d176, mb); cudaMalloc(&d176x, mb); cudaMalloc(&ddev, mb); cudaMalloc(&dtmpl, sizeof(CUtensorMap));
  cudaMemcpy(d32, h32.data(), mb, cudaMemcpyHostToDevice);
  cudaMemcpy(d176, h176.data(), mb, cudaMemcpyHostToDevice);
  cudaMemcpy(d176x, h176x.data(), mb, cudaMemcpyHostToDevice);
  // the template: a valid map of ANOTHER shape on a dummy address; the device fills in address, dims 1-2, strides
  const CUtensorMap tmpl = encode(enc, d_in, 8, 176, 16, 256);
  cudaMemcpy(dtmpl, &tmpl, sizeof(tmpl), cudaMemcpyHostToDevice);
  cudaMemset(ddev, 0, mb);
  k_build_maps<<<n_scans, 32>>>(dtmpl, ddev, d_in, n_scans);
  cudaDeviceSynchronize();
  {
    std::vector<CUtensorMap> back(n_scans);
    cudaMemcpy(back.data(), ddev, mb, cudaMemcpyDeviceToHost);
    int diff = 0;
    for (int s = 0; s < n_scans; s++) { diff += memcmp(&back[s], &h176[s], sizeof(CUtensorMap)) != 0; }
    printf("device-built maps that differ bytewise from host-encoded ones: %d of %d\n", diff, n_scans);
    if (diff) {
      const uint64_t * a = reinterpret_cast<const uint64_t *>(&back[0]), * b = reinterpret_cast<const uint64_t *>(&h176[0]);
      for (int i = 0; i < 16; i++) { printf("  q%02d dev %016llx host %016llx%s\n", i, (unsigned long long)a[i], (unsigned long long)b[i], a[i] != b[i] ? "  <--" : ""); }
    }
  }
  const double items = (double)n_scans * R * B, pts = items * POS;
  printf("%d scans of %d x %d points (%.2f GB), %d SMs, %d warps per CTA, %.0f items of %d positions, delay %d cycles, skew %d %%\n", n_scans, R, W,
         bytes / 1e9, sms, NW, items, POS, delay, skew);
  const char * names[4] = {"TMA boxes 32 B x 32 columns (11 per item)", "TMA boxes 32 B x 176 columns (2 per item)",
                           "TMA boxes 16 B x 176 columns (x,y,z,pad only)", "like 2nd, maps built on the device"};
  unsigned long long bad[4] = {0, 0, 0, 0};
  float ms[4] = {run<0>(d32, n_scans, delay, skew, d_out, sms, bad[0]), run<1>(d176, n_scans, delay, skew, d_out, sms, bad[1]),
                 run<2>(d176x, n_scans, delay, skew, d_out, sms, bad[2]), run<3>(ddev, n_scans, delay, skew, d_out, sms, bad[3])};
  for (int m = 0; m < 4; m++) {
    printf("%-48s %8.3f ms  %7.1f Gpositions/s  %7.1f GB/s of 32-byte points  mismatches %llu\n", names[m], ms[m], pts / ms[m] / 1e6,
           pts * 32 / ms[m] / 1e6, bad[m]);
  }
  return 0;
}
