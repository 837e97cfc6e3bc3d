// gather_probe.cu — what the ring-sector gather costs by itself (no arithmetic), for several ways of moving the
// points of a window into shared memory. Geometry of the bench workload: scans of R = 128 rings x W = 2048 firings,
// 32-byte points in firing order (ring r of a scan = points r, r + R, ...: 4 KB apart); one warp per (ring, sector)
// item stages 32 * K = 352 consecutive positions of its ring, consumes x and the ring word, and moves on; persistent
// grid of one 12-warp CTA per SM, warps of a CTA on neighbouring rings (like k_extract_sectors).
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probes/gather_probe tools/probes/gather_probe.cu
//   run:   tools/probes/gather_probe [scans=256]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifndef PROBE_NW
#define PROBE_NW 12
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

// MODE 0: cp.async 16 B (x,y,z,w)             1: cp.async 16 B + cp.async 4 B (ring word)   [the shipped scheme]
//      2: one 32-byte bulk copy (TMA unit) per point, issued by the point's lane, mbarrier completion
//      3: LDG.256 into registers, then STS.128   4: LDG.128 (x,y,z,w) + LDG.32 (ring word) into registers, STS
template<int MODE>
__global__ void __launch_bounds__(NW * 32, 1) k_gather(const uint8_t * __restrict__ in, int n_scans, unsigned long long * out)
{
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int SLOT = MODE == 2 ? 32 : 16;
  constexpr int PER_WARP = 2 * POS * SLOT + 2 * POS * 4 + 64;   // two item buffers (+ ring words) + barriers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t * base = smem + (size_t)warp * PER_WARP;
  uint8_t * buf[2] = {base, base + POS * SLOT};
  uint32_t * rid[2] = {reinterpret_cast<uint32_t *>(base + 2 * POS * SLOT), reinterpret_cast<uint32_t *>(base + 2 * POS * SLOT + POS * 4)};
  uint64_t * bar = reinterpret_cast<uint64_t *>(base + 2 * POS * SLOT + 2 * POS * 4);
  if (MODE == 2 && lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint32_t n_rings = (uint32_t)n_scans * R;
  const uint32_t n_units = ((n_rings + NW - 1) / NW) * B;
  unsigned long long acc = 0;
  uint32_t phase[2] = {0, 0};
  auto item_addr = [&](uint32_t unit, bool & ok) -> const uint8_t * {
    const uint32_t chunk = unit / B, j = unit % B;
    const uint32_t e = chunk * NW + warp;               // ring index in the batch
    ok = unit < n_units && e < n_rings;
    const uint32_t scan = e / R, r = e % R;
    const uint32_t c0 = j * ((W - POS) / (B - 1));      // first firing of the window (windows overlap a little, like halos)
    return in + ((size_t)scan * W * R + (size_t)c0 * R + r) * 32;
  };
  auto issue = [&](uint32_t unit, int b) {
    bool ok;
    const uint8_t * p = item_addr(unit, ok);
    if (!ok) { return; }
    if (MODE == 2 && lane == 0) { mbar_expect(&bar[b], POS * 32); }
    if (MODE == 2) { __syncwarp(); }
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint8_t * src = p + (size_t)(lane * K + k) * (R * 32);
      if (MODE == 0 || MODE == 1) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(buf[b] + (lane * K + k) * 16)), "l"(src) : "memory");
        if (MODE == 1) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(&rid[b][k * 32 + lane])), "l"(src + 20) : "memory"); }
      } else if (MODE == 2) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                     :: "r"(smem_u32(buf[b] + (lane * K + k) * 32)), "l"(src), "r"(smem_u32(&bar[b])) : "memory");
      }
    }
  };
  if (MODE <= 2) { issue(blockIdx.x, 0); }
  for (uint32_t t = 0; blockIdx.x + t * gridDim.x < n_units; t++) {
    const uint32_t unit = blockIdx.x + t * gridDim.x;
    const int b = t & 1;
    bool ok;
    const uint8_t * p = item_addr(unit, ok);
    if (MODE <= 1) { asm volatile("cp.async.wait_all;" ::: "memory"); __syncwarp(); }
    if (MODE == 2 && ok) { mbar_wait(&bar[b], phase[b]); phase[b] ^= 1; }
    if (NW % 4 == 0) { asm volatile("bar.sync %0, %1;" :: "r"(1 + warp / 4), "r"(128) : "memory"); }   // quads stay aligned, like the real kernel
    if (MODE <= 2) { issue(unit + gridDim.x, b ^ 1); }
    if (!ok) { continue; }
    if (MODE == 3 || MODE == 4) {
      uint32_t v[K][4];
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint8_t * src = p + (size_t)(lane * K + k) * (R * 32);
        if (MODE == 3) {
          uint32_t w3, w4, w6, w7;
          asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[k][0]), "=r"(v[k][1]), "=r"(v[k][2]), "=r"(w3), "=r"(w4), "=r"(v[k][3]), "=r"(w6), "=r"(w7) : "l"(src));
        } else {
          const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src));
          v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z;
          v[k][3] = __ldg(reinterpret_cast<const uint32_t *>(src + 20));
        }
      }
#pragma unroll
      for (int k = 0; k < K; k++) {
        *reinterpret_cast<uint4 *>(buf[b] + (lane * K + k) * 16) = make_uint4(v[k][0], v[k][1], v[k][2], 0);
        acc += v[k][3] & 0xFFFF;
      }
      __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
      acc += *reinterpret_cast<const uint32_t *>(buf[b] + (lane * K + k) * SLOT);
      if (MODE == 1) { acc += rid[b][k * 32 + lane] & 0xFFFF; }
      if (MODE == 2) { acc += *reinterpret_cast<const uint32_t *>(buf[b] + (lane * K + k) * SLOT + 20) & 0xFFFF; }
    }
    __syncwarp();
  }
  if (MODE <= 1) { asm volatile("cp.async.wait_all;" ::: "memory"); }
  if (acc == 0x123456789ull) { out[0] = acc; }
}

template<int MODE>
float run(const uint8_t * d_in, int n_scans, unsigned long long * d_out, int sms)
{
  constexpr int SLOT = MODE == 2 ? 32 : 16;
  const size_t smem = (size_t)NW * (2 * POS * SLOT + 2 * POS * 4 + 64);
  cudaFuncSetAttribute(k_gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) { k_gather<MODE><<<sms, NW * 32, smem>>>(d_in, n_scans, d_out); }
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) { k_gather<MODE><<<sms, NW * 32, smem>>>(d_in, n_scans, d_out); }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(err)); return -1.f; }
  return ms / 5;
}

int main(int argc, char ** argv)
{
  const int n_scans = argc > 1 ? atoi(argv[1]) : 256;
  const size_t bytes = (size_t)n_scans * R * W * 32;
  uint8_t * d_in; unsigned long long * d_out;
  cudaMalloc(&d_in, bytes); cudaMalloc(&d_out, 8);
  cudaMemset(d_in, 1, bytes);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const double items = (double)n_scans * R * B, pts = items * POS;
  const char * names[5] = {"cp.async 16 B", "cp.async 16 B + 4 B (shipped)", "bulk copy 32 B per point (TMA unit)", "LDG.256 -> STS.128", "LDG.128 + LDG.32 -> STS.128"};
  float ms[5] = {run<0>(d_in, n_scans, d_out, sms), run<1>(d_in, n_scans, d_out, sms), run<2>(d_in, n_scans, d_out, sms),
                 run<3>(d_in, n_scans, d_out, sms), run<4>(d_in, n_scans, d_out, sms)};
  printf("%d scans of %d x %d points (%.2f GB), %d SMs, %.0f items of %d positions\n", n_scans, R, W, bytes / 1e9, sms, items, POS);
  for (int m = 0; m < 5; m++) {
    printf("%-40s %8.3f ms  %7.1f Gpositions/s  %7.1f GB/s of 32-byte points\n", names[m], ms[m], pts / ms[m] / 1e6, pts * 32 / ms[m] / 1e6);
  }
  return 0;
}
