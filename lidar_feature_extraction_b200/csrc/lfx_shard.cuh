// lfx_shard.cuh — the path's one exchange between GPUs: every rank's per-scan feature counts reach
// every rank, which turns them into frame-ordered global offsets.
//
// NCCL's all-gather does that with a kernel of its own, enqueued behind a persistent extraction kernel that holds every
// SM: 28-89 us per step on 2-8 B200 (profiles/r01s_scaling.md). Here the counts are PUSHED instead: a one-CTA kernel
// at the tail of the batch stores this rank's 10 KB block into every peer's symmetric receive buffer through peer-mapped
// memory (NVLink; CUDA IPC handles between processes, plain peer access inside one process) and then raises a per-rank
// epoch flag there. The consumer side (k_shard_scan) runs one batch later, when the flags have long arrived: it checks
// them (bounded spin), puts the blocks into frame order and scans them. NCCL is still what sets the group up (unique
// id rendezvous, exchange of the IPC handles) and is the fallback exchange when peer mapping is not available.
#ifndef LFX_SHARD_CUH_
#define LFX_SHARD_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace lfxk
{

constexpr int SHARD_MAX_WORLD = 64;
constexpr int SHARD_THREADS = 1024;
// receive slots per rank, used round robin by epoch. Three, not two: k_shard_scan_push stores the counts of epoch e + 1 into
// the peers BEFORE it waits for their flags of epoch e, i.e. possibly while a peer one batch behind still scans its slot of
// epoch e - 1; slot (e + 2) % 3 = (e - 1) % 3 is only written after this rank has seen that peer's flag of epoch e + 1,
// which the peer raises after that scan.
constexpr uint32_t SHARD_SLOTS = 3;

// layout of one receive slot: [world][width][2] counts, then [world] epoch flags (all uint32)
// (an even number of words: rows are read as 8-byte pairs in either slot)
__host__ __device__ inline size_t shard_slot_words(int world, size_t width) { return ((size_t)world * width * 2 + (size_t)world + 1) & ~(size_t)1; }

struct ShardPeers { uint32_t * slot[SHARD_MAX_WORLD]; };   // peer p's receive slot of the current epoch (epoch % SHARD_SLOTS), as mapped here

// this rank's counts -> block `rank` of every peer's slot, then the flag (release at system scope)
struct ShardPushArgs { const uint32_t * counts; uint32_t n_local; ShardPeers peers; int rank, world; uint32_t width, epoch; };

// (two halves: the stores travel over NVLink while the kernel does something else - the scan of the previous exchange -
//  and the system-scope fence in front of the flags then finds them delivered)
__device__ __forceinline__ void shard_push_data(const uint32_t * __restrict__ counts, uint32_t n_local, const ShardPeers & peers, int rank, int world,
                                                uint32_t width)
{
  const uint32_t words = 2u * n_local;
  for (int p = 0; p < world; p++) {
    uint32_t * dst = peers.slot[p] + (size_t)rank * width * 2;
    for (uint32_t i = threadIdx.x; i < words; i += SHARD_THREADS) { dst[i] = counts[i]; }
  }
}
__device__ __forceinline__ void shard_push_flags(const ShardPeers & peers, int rank, int world, uint32_t width, uint32_t epoch)
{
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    uint32_t * flag = peers.slot[threadIdx.x] + (size_t)world * width * 2 + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(epoch) : "memory");
  }
}
__device__ __forceinline__ void shard_push_body(const uint32_t * __restrict__ counts, uint32_t n_local, const ShardPeers & peers, int rank, int world,
                                                uint32_t width, uint32_t epoch)
{
  shard_push_data(counts, n_local, peers, rank, world, width);
  shard_push_flags(peers, rank, world, width, epoch);
}

struct ShardScanArgs
{
  const uint32_t * slot;       // this rank's receive slot: blocks of all ranks (+ flags when pushed)
  uint32_t * counts_all;       // [n_frames][2] in frame order
  unsigned long long * offsets_all;   // [n_frames + 1][2]
  uint32_t * status;           // 0 ok, 1: a peer's flag did not arrive in time
  unsigned long long n_frames;
  int world;
  uint32_t width;
  uint32_t epoch;              // 0: no flags to wait for (NCCL all-gather filled the slot)
  unsigned long long timeout_ns;
};

// rank g owns frames [g F / G, (g + 1) F / G): its block holds (g + 1) F / G - g F / G rows
__device__ __forceinline__ unsigned long long shard_first(unsigned long long F, int g, int G) { return ((unsigned long long)g * F) / (unsigned long long)G; }

__device__ __forceinline__ void shard_scan_body(const ShardScanArgs & a)
{
  __shared__ unsigned long long s_e[64], s_s[64];
  __shared__ unsigned long long s_first[SHARD_MAX_WORLD + 1];   // first frame of every rank's block (one division each, not per frame)
  __shared__ int s_late;
  const int tid = threadIdx.x;
  if (tid == 0) { s_late = 0; }
  if (tid <= a.world) { s_first[tid] = shard_first(a.n_frames, tid, a.world); }
  __syncthreads();
  if (a.epoch != 0 && tid < a.world) {
    const uint32_t * flag = a.slot + (size_t)a.world * a.width * 2 + tid;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v == a.epoch) { break; }
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > a.timeout_ns) { s_late = 1; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (s_late) { if (tid == 0) { *a.status = 1u; } return; }
  // frame f lives in block g = owner(f) at row f - first(g). Frames are walked in chunks of 4 x 1024 with a running
  // carry: four consecutive frames per thread, warp scan by shuffles, warp totals scanned by warp 0.
  constexpr int PER = 4;
  const int lane = tid & 31, warp = tid >> 5;
  unsigned long long carry_e = 0, carry_s = 0;
  for (unsigned long long f0 = 0; f0 < a.n_frames; f0 += (unsigned long long)PER * SHARD_THREADS) {
    uint32_t ne[PER], ns[PER];
    unsigned long long te = 0, ts = 0;
    // owner of the thread's first frame: the largest g with first(g) <= f (binary search over the table); the next
    // three frames are its or a neighbour's
    int g = 0;
    {
      const unsigned long long f = min(f0 + (unsigned long long)tid * PER, a.n_frames - 1);
      int lo = 0, hi = a.world - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_first[mid] <= f) { lo = mid; } else { hi = mid - 1; }
      }
      g = lo;
    }
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const unsigned long long f = f0 + (unsigned long long)tid * PER + u;
      ne[u] = 0; ns[u] = 0;
      if (f < a.n_frames) {
        while (g + 1 < a.world && s_first[g + 1] <= f) { g++; }
        const unsigned long long row = f - s_first[g];
        const uint2 v = *reinterpret_cast<const uint2 *>(a.slot + ((size_t)g * a.width + row) * 2);
        ne[u] = v.x; ns[u] = v.y;
        *reinterpret_cast<uint2 *>(a.counts_all + 2 * f) = v;
      }
      te += ne[u]; ts += ns[u];
    }
    unsigned long long ie = te, is = ts;   // inclusive scan of the threads' totals inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long ve = __shfl_up_sync(0xFFFFFFFFu, ie, o), vs = __shfl_up_sync(0xFFFFFFFFu, is, o);
      if (lane >= o) { ie += ve; is += vs; }
    }
    __syncthreads();   // the previous chunk's warp totals have been read
    if (lane == 31) { s_e[warp] = ie; s_s[warp] = is; }
    __syncthreads();
    if (warp == 0) {
      unsigned long long we = s_e[lane], ws = s_s[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ve = __shfl_up_sync(0xFFFFFFFFu, we, o), vs = __shfl_up_sync(0xFFFFFFFFu, ws, o);
        if (lane >= o) { we += ve; ws += vs; }
      }
      s_e[32 + lane] = we; s_s[32 + lane] = ws;   // inclusive over the warps
    }
    __syncthreads();
    unsigned long long be = carry_e + (warp ? s_e[32 + warp - 1] : 0ull) + ie - te;
    unsigned long long bs = carry_s + (warp ? s_s[32 + warp - 1] : 0ull) + is - ts;
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const unsigned long long f = f0 + (unsigned long long)tid * PER + u;
      if (f < a.n_frames) { a.offsets_all[2 * f] = be; a.offsets_all[2 * f + 1] = bs; }
      be += ne[u]; bs += ns[u];
    }
    carry_e += s_e[63]; carry_s += s_s[63];
  }
  if (tid == 0) {
    a.offsets_all[2 * a.n_frames] = carry_e;
    a.offsets_all[2 * a.n_frames + 1] = carry_s;
    *a.status = 0u;
  }
}

__global__ void __launch_bounds__(SHARD_THREADS) k_shard_scan(const ShardScanArgs a) { shard_scan_body(a); }
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_push(const ShardPushArgs p)
{
  shard_push_body(p.counts, p.n_local, p.peers, p.rank, p.world, p.width, p.epoch);
}
// the steady state of a sequence of batches: the consumer side of the previous exchange and the producer side of this
// one in ONE launch at the tail of the batch
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_scan_push(const ShardScanArgs a, const ShardPushArgs p)
{
  shard_push_data(p.counts, p.n_local, p.peers, p.rank, p.world, p.width);   // (slot epoch % 3: the scan below reads the previous epoch's)
  shard_scan_body(a);
  __syncthreads();
  shard_push_flags(p.peers, p.rank, p.world, p.width, p.epoch);
}

}  // namespace lfxk
#endif  // LFX_SHARD_CUH_
