// Gradient-bucket all-reduce fused with the AMP unscale, the 1/world average and the overflow check,
// issued as plain loads/stores (or multimem instructions) on NVLink-mapped peer memory -- sm_100a.
//
// Replaces, per bucket, the NCCL all-reduce that DDP's C++ Reducer fires during backward, the
// `div_(world)` pre-pass, the bucket copy-in/copy-out, and Apex's `amp_C.multi_tensor_scale` +
// inf/nan scan after backward (BASELINE.json north star).  The bucket is a
// slice of the flat gradient arena, which lives in symmetric memory: every rank holds a mapping of
// every other rank's arena (peer_bufs) and, when NVSwitch multicast is available, one multicast
// mapping (mc_buf).  The reduction is in place on all ranks.
//
// Variants (picked per bucket size):
//   one-shot   (<= 256 KiB) every rank loads the whole bucket from all peers, adds in registers, waits
//              for everyone to finish reading, then overwrites its own copy.  2 barriers, 1 hop.
//   two-shot   rank r owns slice r: reduce-scatter (load slice r from all peers) + all-gather (store the
//              scaled sum into slice r of all peers).  Wire bytes per GPU: 2 (n-1)/n S.
//   multimem   as two-shot, but the switch does the add (multimem.ld_reduce) and the broadcast
//              (multimem.st): one load and one store per element regardless of n.
//
// Cross-GPU ordering: a CTA-level barrier is built from system-scope CAS on per-(channel, block, peer)
// words of the peers' signal pads (release on signal, acquire on wait; the surrounding bar.sync
// extends the ordering to the whole CTA).  Flags toggle 0 -> 1 -> 0, so the protocol needs no
// sequence numbers and is CUDA-graph replay safe.  Waits are bounded: a missing peer traps with a
// message instead of hanging the GPU.
#include "../include/ddpt_comm.h"
#include "../include/ddpt_common.cuh"
#include <cstdio>
#include <stdexcept>

namespace ddpt {
namespace {

constexpr int kThreads = 512;
constexpr long long kOneShotMaxElems = 64 * 1024;      // 256 KiB of fp32

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Wall-clock bounded spin: `deadline` is armed on the first failed probe.
__device__ __forceinline__ void spin_guard(unsigned long long& deadline, unsigned long long timeout_ns,
                                           const char* what, int rank) {
  const unsigned long long now = global_ns();
  if (deadline == 0) deadline = now + timeout_ns;
  else if (now > deadline) {
    printf("ddpt fused_allreduce: rank %d timed out in %s (block %d)\n", rank, what, blockIdx.x);
    __trap();
  }
}

// All CTAs with the same blockIdx.x on all ranks meet here.
__device__ __forceinline__ void cross_gpu_barrier(const AllReduceArgs& a) {
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    const int slot = (a.channel * kCommMaxBlocks + blockIdx.x) * kMaxRanks;
    unsigned long long deadline = 0;
    const unsigned long long timeout_ns = (unsigned long long)a.timeout_ms * 1000000ull;
    uint32_t* remote = a.peer_signals[peer] + slot + a.rank;      // "rank a.rank has arrived" on the peer
    while (cas_release_sys(remote, 0u, 1u) != 0u) spin_guard(deadline, timeout_ns, "signal", a.rank);
    uint32_t* local = a.peer_signals[a.rank] + slot + peer;        // wait for the peer's arrival here
    while (cas_acquire_sys(local, 1u, 0u) != 1u) spin_guard(deadline, timeout_ns, "wait", a.rank);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_f4(float* p, const float4& v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_f4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ bool finite4(const float4& v) {
  return is_finite_f32(v.x) && is_finite_f32(v.y) && is_finite_f32(v.z) && is_finite_f32(v.w);
}

__device__ __forceinline__ float scale_of(const AllReduceArgs& a) {
  const float inv = a.scaler_state != nullptr ? a.scaler_state[kInvScale] : 1.0f;
  return inv / (float)a.world;
}

// Overflow bookkeeping: the owner of a slice raises the channel's word on EVERY rank; after the closing
// barrier each rank folds its word into its local found_inf and clears it.
__device__ __forceinline__ void publish_overflow(const AllReduceArgs& a, bool bad) {
  if (!__syncthreads_or(bad)) return;
  if ((int)threadIdx.x < a.world)
    a.peer_signals[threadIdx.x][kCommChannels * kCommMaxBlocks * kMaxRanks + a.channel] = 1u;
  __threadfence_system();
}
__device__ __forceinline__ void collect_overflow(const AllReduceArgs& a) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t* word = a.peer_signals[a.rank] + kCommChannels * kCommMaxBlocks * kMaxRanks + a.channel;
    if (*reinterpret_cast<volatile uint32_t*>(word) != 0u) {
      if (a.found_inf != nullptr) *a.found_inf = 1.0f;
      *word = 0u;
    }
  }
}

template <int ALGO>
__global__ void __launch_bounds__(kThreads)
fused_allreduce_kernel(AllReduceArgs a) {
  const float scale = scale_of(a);
  bool bad = false;
  cross_gpu_barrier(a);                                   // everyone's gradients for this bucket are complete
  if (ALGO == kAlgoOneShot) {
    const long long nvec = a.numel >> 2;
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;   // at most one vector per thread
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nvec) {
#pragma unroll 1
      for (int r = 0; r < a.world; ++r) {
        const float4 v = ld_f4(a.peer_bufs[(a.rank + r) % a.world] + a.offset + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
      bad = !finite4(acc);
    }
    cross_gpu_barrier(a);                                 // all ranks have finished reading
    if (i < nvec) st_f4(a.peer_bufs[a.rank] + a.offset + 4 * i, acc);
    if (__syncthreads_or(bad) && threadIdx.x == 0 && a.found_inf != nullptr) *a.found_inf = 1.0f;
    return;                                               // every rank saw the same sums: flags agree
  }
  const long long slice = a.numel / a.world;              // elements owned by this rank
  const long long nvec = slice >> 2;
  const long long base = a.offset + (long long)a.rank * slice;
  // kUnroll independent 16-byte vectors per thread are in flight at once: NVLink round trips are ~2 us, so
  // bandwidth is bytes-in-flight limited (blocks x 512 threads x kUnroll x 16 B x world loads).
  constexpr int kUnroll = 8;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i0 = (long long)blockIdx.x * kThreads + threadIdx.x; i0 < nvec; i0 += stride * kUnroll) {
    float4 acc[kUnroll];
    if (ALGO == kAlgoMultimem) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i < nvec) acc[u] = mc_ld_reduce_f4(a.mc_buf + base + 4 * i);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int r = 0; r < a.world; ++r) {
        const float* src = a.peer_bufs[(a.rank + r) % a.world] + base;
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec) v[u] = ld_f4(src + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (i0 + u * stride < nvec) {
            acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (i0 + u * stride < nvec) {
        acc[u].x *= scale; acc[u].y *= scale; acc[u].z *= scale; acc[u].w *= scale;
        bad |= !finite4(acc[u]);
      }
    }
    if (ALGO == kAlgoMultimem) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i < nvec) mc_st_f4(a.mc_buf + base + 4 * i, acc[u]);
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < a.world; ++r) {
        float* dst = a.peer_bufs[(a.rank + r) % a.world] + base;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec) st_f4(dst + 4 * i, acc[u]);
        }
      }
    }
  }
  publish_overflow(a, bad);
  cross_gpu_barrier(a);                                   // all slices (and overflow words) have landed everywhere
  collect_overflow(a);
}

__global__ void __launch_bounds__(kThreads)
symm_broadcast_kernel(AllReduceArgs a, int root) {
  cross_gpu_barrier(a);
  if (a.rank == root) {
    const long long nvec = a.numel >> 2;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (long long)gridDim.x * kThreads) {
      const float4 v = ld_f4(a.peer_bufs[root] + a.offset + 4 * i);
      for (int r = 0; r < a.world; ++r)
        if (r != root) st_f4(a.peer_bufs[r] + a.offset + 4 * i, v);
    }
  }
  cross_gpu_barrier(a);
}

}  // namespace

void launch_fused_allreduce(const AllReduceArgs& args, cudaStream_t stream) {
  if (args.world < 1 || args.world > kMaxRanks) throw std::runtime_error("fused_allreduce: world must be in [1, 8]");
  if (args.numel % (4 * args.world) != 0 || args.offset % 4 != 0)
    throw std::runtime_error("fused_allreduce: bucket must be a multiple of 4*world elements, 16-byte aligned");
  if (args.channel < 0 || args.channel >= kCommChannels) throw std::runtime_error("fused_allreduce: bad channel");
  int algo = args.algo;
  if (algo == kAlgoAuto) {
    if (args.numel <= kOneShotMaxElems) algo = kAlgoOneShot;
    else algo = args.mc_buf != nullptr ? kAlgoMultimem : kAlgoTwoShot;
  }
  if (algo == kAlgoMultimem && args.mc_buf == nullptr) throw std::runtime_error("fused_allreduce: no multicast mapping");
  AllReduceArgs a = args;
  if (a.timeout_ms <= 0) a.timeout_ms = 30000;
  int blocks;
  if (algo == kAlgoOneShot) {
    if (args.numel > kOneShotMaxElems * 1) throw std::runtime_error("fused_allreduce: bucket too large for one-shot");
    blocks = (int)(((args.numel >> 2) + kThreads - 1) / kThreads);
    if (blocks < 1) blocks = 1;
    if (blocks > kCommMaxBlocks) throw std::runtime_error("fused_allreduce: one-shot needs <= 32 blocks");
    fused_allreduce_kernel<kAlgoOneShot><<<blocks, kThreads, 0, stream>>>(a);
  } else {
    const long long vec_per_rank = (args.numel / args.world) >> 2;
    long long want = (vec_per_rank + kThreads - 1) / kThreads;
    blocks = args.blocks > 0 ? args.blocks : 32;
    if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
    if (want < blocks) blocks = (int)(want > 0 ? want : 1);
    if (algo == kAlgoMultimem) fused_allreduce_kernel<kAlgoMultimem><<<blocks, kThreads, 0, stream>>>(a);
    else fused_allreduce_kernel<kAlgoTwoShot><<<blocks, kThreads, 0, stream>>>(a);
  }
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_symm_broadcast(const AllReduceArgs& args, int root, cudaStream_t stream) {
  if (args.numel % 4 != 0) throw std::runtime_error("symm_broadcast: numel must be a multiple of 4");
  long long want = ((args.numel >> 2) + kThreads - 1) / kThreads;
  int blocks = (int)(want < 8 ? (want > 0 ? want : 1) : 8);
  AllReduceArgs a = args;
  if (a.timeout_ms <= 0) a.timeout_ms = 30000;
  symm_broadcast_kernel<<<blocks, kThreads, 0, stream>>>(a, root);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
