// Gradient-bucket all-reduce fused with the wire-format cast (fp32 -> bf16 on send, fp32 accumulate), the AMP
// unscale, the 1/world average and the overflow check, issued as plain loads/stores (or multimem instructions)
// on NVLink-mapped peer memory -- sm_100a.
//
// Replaces, per bucket, the NCCL all-reduce that DDP's C++ Reducer fires during backward, the `div_(world)`
// pre-pass, the bucket copy-in/copy-out, and Apex's `amp_C.multi_tensor_scale` + inf/nan scan after backward
// (BASELINE.json north star; reference call sites utils/process.py:26-27,
// utils/train.py:61).  The bucket is a slice of the flat gradient arena, which lives in symmetric memory:
// every rank holds a mapping of every other rank's arena (peer_bufs), of a scratch region (peer_scratch) and,
// when NVSwitch multicast is available, one multicast mapping of each (mc_buf / mc_scratch).  The reduction
// is in place on all ranks.
//
// Variants (picked per bucket size), each with an fp32 or a bf16 wire format:
//   one-shot   (<= 256 KiB) PUSH: every rank writes its (optionally bf16-converted) bucket into slot [rank] of
//              every peer's scratch (one multimem.st per vector, or world peer stores), ONE barrier, then
//              every rank sums the world slots locally in rank order (bit-identical results everywhere) and
//              writes the scaled fp32 result into its own arena.  The slots are double-buffered by epoch
//              parity, so no closing barrier is needed: a rank can run at most one launch ahead of a peer.
//   two-shot   rank r owns slice r: reduce-scatter (load slice r from all peers) + all-gather (store the
//              scaled sum into slice r of all peers).  Wire bytes per GPU: 2 (n-1)/n S.
//   multimem   as two-shot, but the switch does the add (multimem.ld_reduce, fp32 accumulation) and the
//              broadcast (multimem.st): one load and one store per element regardless of n.
//   bf16 wire (two-shot / multimem): the kernel first casts its own bucket into the 16-bit staging region
//              (local pass), reduces and gathers over the staging (half the NVLink bytes), and finally
//              expands the averaged staging into the fp32 arena (local pass).
//
// Cross-GPU ordering: a CTA-level barrier built from one `st.release.sys` per peer (arrive: fire and forget,
// no round trip) and one `ld.acquire.sys` poll per peer on the local signal pad.  Flags carry a monotonic
// epoch that lives in device memory ((epoch << 1) | payload), so nothing is ever reset, the protocol is
// CUDA-graph replay safe, and the closing barrier's payload bit carries each owner's overflow verdict to
// every rank without a separate (racy) word.  Waits are wall-clock bounded: a missing peer traps with a
// message instead of hanging the GPU.
#include "../include/ddpt_comm.h"
#include "../include/ddpt_common.cuh"
#include <cstdio>
#include <stdexcept>

namespace ddpt {
namespace {

// 256 threads x <= 64 registers = 16 Ki registers per CTA: a communication CTA then CO-RESIDES with a persistent GEMM CTA
// (320 threads x 144 registers = 45 Ki of the SM's 64 Ki).  The first version ran 512 x 98 registers: every SM that hosted
// one could not accept the GEMM CTA scheduled next, and because the GEMM kernels walk static tile lists, a CTA that
// starts 60 us late finishes 60 us late -- the overlapped all-reduce stretched the backward it was hiding under.
constexpr int kThreads = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t val) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(val) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
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

__device__ __forceinline__ uint32_t* epoch_word(const AllReduceArgs& a) {
  return a.peer_signals[a.rank] + kFlagWords + a.channel * kCommMaxBlocks + blockIdx.x;
}
// Epoch of this launch for (channel, block): previous launch's value + 1.  The word is only written by the same
// block index of the previous launch on this channel, which precedes this one in stream order.
__device__ __forceinline__ uint32_t epoch_begin(const AllReduceArgs& a) {
  return *reinterpret_cast<volatile uint32_t*>(epoch_word(a)) + 1u;
}
__device__ __forceinline__ void epoch_end(const AllReduceArgs& a, uint32_t e) {
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(epoch_word(a)) = e;
}

// All CTAs with the same blockIdx.x on all ranks meet here.  Returns the OR of the peers' payload bits.
__device__ __forceinline__ bool cross_gpu_barrier(const AllReduceArgs& a, int phase, uint32_t e, bool payload) {
  __syncthreads();
  uint32_t got = 0;
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    const int slot = ((a.channel * kCommMaxBlocks + blockIdx.x) * 2 + phase) * kMaxRanks;
    st_release_sys(a.peer_signals[peer] + slot + a.rank, (e << 1) | (payload ? 1u : 0u));   // "rank a.rank arrived"
    const uint32_t* local = a.peer_signals[a.rank] + slot + peer;                             // the peer's arrival
    unsigned long long deadline = 0;
    const unsigned long long timeout_ns = (unsigned long long)a.timeout_ms * 1000000ull;
    uint32_t v;
    while (((v = ld_acquire_sys(local)) >> 1) < e) spin_guard(deadline, timeout_ns, "barrier", a.rank);
    got = v & 1u;
  }
  return __syncthreads_or((int)got) != 0;
}

// ---- 16-byte accessors ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_f4(float* p, const float4& v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_u4(void* p, const uint4& v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// 8 bf16 values summed by the switch with fp32 accumulation, returned as packed bf16x2
__device__ __forceinline__ uint4 mc_ld_reduce_bf16x8(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_f4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st_u4(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}

__device__ __forceinline__ bool finite4(const float4& v) {
  return is_finite_f32(v.x) && is_finite_f32(v.y) && is_finite_f32(v.z) && is_finite_f32(v.w);
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void mul4(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }
__device__ __forceinline__ uint4 pack8(const float4& lo, const float4& hi) {
  return make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
}
__device__ __forceinline__ void unpack8(const uint4& v, float4& lo, float4& hi) {
  lo = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                   __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
  hi = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u),
                   __uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u));
}

__device__ __forceinline__ float scale_of(const AllReduceArgs& a) {
  const float inv = a.scaler_state != nullptr ? a.scaler_state[kInvScale] : 1.0f;
  return inv / (float)a.world;
}

__device__ __forceinline__ long long push_offset(int channel, uint32_t parity, int src) {
  return (((long long)channel * 2 + parity) * kMaxRanks + src) * kPushSlotBytes;
}

// ---- one-shot (push) ----------------------------------------------------------------------------------------------
template <int WIRE, bool MC>
__global__ void __launch_bounds__(kThreads, 4)
oneshot_push_kernel(AllReduceArgs a) {
  const float scale = scale_of(a);
  const uint32_t e = epoch_begin(a);
  const uint32_t parity = e & 1u;
  constexpr int kElemsPerVec = WIRE == kWireBF16 ? 8 : 4;
  const long long nvec = a.numel / kElemsPerVec;
  const long long stride = (long long)gridDim.x * kThreads;
  const float* mine = a.peer_bufs[a.rank] + a.offset;
  const long long my_slot = push_offset(a.channel, parity, a.rank);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    uint4 bits;
    if (WIRE == kWireBF16) {
      const float4 lo = *reinterpret_cast<const float4*>(mine + 8 * i);
      const float4 hi = *reinterpret_cast<const float4*>(mine + 8 * i + 4);
      bits = pack8(lo, hi);                                        // the cast rides on the send
    } else {
      bits = *reinterpret_cast<const uint4*>(mine + 4 * i);
    }
    if (MC) {
      mc_st_u4(a.mc_scratch + my_slot + 16 * i, bits);
    } else {
#pragma unroll 1
      for (int r = 0; r < a.world; ++r) st_u4(a.peer_scratch[(a.rank + r) % a.world] + my_slot + 16 * i, bits);
    }
  }
  cross_gpu_barrier(a, 0, e, false);                               // every rank's bucket has landed in my scratch
  bool bad = false;
  float* out = a.peer_bufs[a.rank] + a.offset;
  const unsigned char* slots = a.peer_scratch[a.rank] + push_offset(a.channel, parity, 0);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll 1
    for (int r = 0; r < a.world; ++r) {                            // fixed rank order: identical sums on every rank
      const uint4 bits = ld_u4(slots + (long long)r * kPushSlotBytes + 16 * i);
      if (WIRE == kWireBF16) {
        float4 lo, hi;
        unpack8(bits, lo, hi);
        add4(acc0, lo); add4(acc1, hi);
      } else {
        add4(acc0, make_float4(__uint_as_float(bits.x), __uint_as_float(bits.y), __uint_as_float(bits.z),
                               __uint_as_float(bits.w)));
      }
    }
    mul4(acc0, scale);
    bad |= !finite4(acc0);
    if (WIRE == kWireBF16) {
      mul4(acc1, scale);
      bad |= !finite4(acc1);
      *reinterpret_cast<float4*>(out + 8 * i) = acc0;
      *reinterpret_cast<float4*>(out + 8 * i + 4) = acc1;
    } else {
      *reinterpret_cast<float4*>(out + 4 * i) = acc0;
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0 && a.found_inf != nullptr) *a.found_inf = 1.0f;
  epoch_end(a, e);
}

// ---- two-shot (peer loads/stores or multimem) -----------------------------------------------------------------------
template <int WIRE, bool MC>
__global__ void __launch_bounds__(kThreads, 4)
twoshot_kernel(AllReduceArgs a) {
  const float scale = scale_of(a);
  const uint32_t e = epoch_begin(a);
  const long long stride = (long long)gridDim.x * kThreads;
  const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
  bool bad = false;
  // kUnroll independent 16-byte vectors per thread are in flight at once: NVLink round trips are ~2 us, so
  // bandwidth is bytes-in-flight limited (blocks x 512 threads x kUnroll x 16 B x world loads).
  constexpr int kUnroll = MC ? 8 : (WIRE == kWireBF16 ? 2 : 4);     // peer-load variants hold fp32 accumulators (64-register budget)

  const long long slice = a.numel / a.world;              // elements owned by each rank
  if (WIRE == kWireBF16) {
    // phase A: cast my bucket into my 16-bit staging (local read 4 B/elem, local write 2 B/elem).  The barriers pair
    // CTA b of every rank with CTA b of every other rank and nothing else, so CTA b must produce exactly the staging
    // vectors that the owners' CTA b will read in phase B: for every slice, the same (tid + k * stride) index set that
    // phase B walks.  (A first version cast the bucket with a flat grid-stride loop: a peer's CTA could then read
    // vectors that a *different* local CTA had not written yet -- invisible in eager runs, wrong gradients by a factor
    // of world x loss_scale under CUDA-graph replay, where co-running kernels delay CTAs.)
    const long long nvec_slice = slice >> 3;
#pragma unroll 1
    for (int r = 0; r < a.world; ++r) {
      const float* mine = a.peer_bufs[a.rank] + a.offset + (long long)r * slice;
      unsigned char* stage = a.peer_scratch[a.rank] + kPushBytes + 2 * (a.offset + (long long)r * slice);
      for (long long i0 = tid; i0 < nvec_slice; i0 += stride * 4) {
        float4 lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec_slice) {
            lo[u] = ld_nc_f4(reinterpret_cast<const float4*>(mine + 8 * i));
            hi[u] = ld_nc_f4(reinterpret_cast<const float4*>(mine + 8 * i + 4));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec_slice) *reinterpret_cast<uint4*>(stage + 16 * i) = pack8(lo[u], hi[u]);
        }
      }
    }
  }
  cross_gpu_barrier(a, 0, e, false);                      // everyone's gradients (staging) for this bucket are complete

  if (WIRE == kWireBF16) {
    const long long nvec = slice >> 3;
    const long long base = kPushBytes + 2 * (a.offset + (long long)a.rank * slice);      // byte offset in scratch
    for (long long i0 = tid; i0 < nvec; i0 += stride * kUnroll) {
      uint4 res[kUnroll];
      if (MC) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec) res[u] = mc_ld_reduce_bf16x8(a.mc_scratch + base + 16 * i);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (i0 + u * stride < nvec) {
            float4 lo, hi;
            unpack8(res[u], lo, hi);
            mul4(lo, scale); mul4(hi, scale);
            bad |= !(finite4(lo) && finite4(hi));
            res[u] = pack8(lo, hi);
          }
        }
      } else {
        float4 lo[kUnroll], hi[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) lo[u] = hi[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r = 0; r < a.world; ++r) {
          const unsigned char* src = a.peer_scratch[(a.rank + r) % a.world] + base;
          uint4 v[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + u * stride;
            if (i < nvec) v[u] = ld_u4(src + 16 * i);
          }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            if (i0 + u * stride < nvec) {
              float4 l, h;
              unpack8(v[u], l, h);
              add4(lo[u], l); add4(hi[u], h);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if (i0 + u * stride < nvec) {
            mul4(lo[u], scale); mul4(hi[u], scale);
            bad |= !(finite4(lo[u]) && finite4(hi[u]));
            res[u] = pack8(lo[u], hi[u]);
          }
        }
      }
      if (MC) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec) mc_st_u4(a.mc_scratch + base + 16 * i, res[u]);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < a.world; ++r) {
          unsigned char* dst = a.peer_scratch[(a.rank + r) % a.world] + base;
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + u * stride;
            if (i < nvec) st_u4(dst + 16 * i, res[u]);
          }
        }
      }
    }
  } else {
    const long long nvec = slice >> 2;
    const long long base = a.offset + (long long)a.rank * slice;
    for (long long i0 = tid; i0 < nvec; i0 += stride * kUnroll) {
      float4 acc[kUnroll];
      if (MC) {
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
          for (int u = 0; u < kUnroll; ++u)
            if (i0 + u * stride < nvec) add4(acc[u], v[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (i0 + u * stride < nvec) {
          mul4(acc[u], scale);
          bad |= !finite4(acc[u]);
        }
      }
      if (MC) {
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
  }
  // closing barrier: all slices have landed everywhere; its payload bit tells every rank whether the owner of
  // (slice, block) saw inf/nan, so all ranks raise found_inf together (no separate flag word, no reset)
  const bool any_bad = cross_gpu_barrier(a, 1, e, __syncthreads_or(bad) != 0);
  if (any_bad && threadIdx.x == 0 && a.found_inf != nullptr) *a.found_inf = 1.0f;

  if (WIRE == kWireBF16) {
    // phase C: expand the averaged 16-bit staging into my fp32 arena (local).  Same CTA-consistent index sets as phase
    // A: slice r's vectors (tid + k * stride) were stored by owner r's CTA b, the one this CTA just met at the barrier.
    const long long nvec_slice = slice >> 3;
#pragma unroll 1
    for (int r = 0; r < a.world; ++r) {
      float* mine = a.peer_bufs[a.rank] + a.offset + (long long)r * slice;
      const unsigned char* stage = a.peer_scratch[a.rank] + kPushBytes + 2 * (a.offset + (long long)r * slice);
      for (long long i0 = tid; i0 < nvec_slice; i0 += stride * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec_slice) v[u] = ld_u4(stage + 16 * i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long i = i0 + u * stride;
          if (i < nvec_slice) {
            float4 lo, hi;
            unpack8(v[u], lo, hi);
            *reinterpret_cast<float4*>(mine + 8 * i) = lo;
            *reinterpret_cast<float4*>(mine + 8 * i + 4) = hi;
          }
        }
      }
    }
  }
  epoch_end(a, e);
}

// root's copy of [offset, offset + numel) replaces everyone's.  Opening barrier: every rank's earlier work on the
// region (readers of the old values) is complete; closing barrier: the new values have landed everywhere.
template <bool MC>
__global__ void __launch_bounds__(kThreads)
symm_broadcast_kernel(AllReduceArgs a, int root) {
  const uint32_t e = epoch_begin(a);
  cross_gpu_barrier(a, 0, e, false);
  if (a.rank == root) {
    const long long nvec = a.numel >> 2;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (long long)gridDim.x * kThreads) {
      const float4 v = *reinterpret_cast<const float4*>(a.peer_bufs[root] + a.offset + 4 * i);
      if (MC) {
        mc_st_f4(a.mc_buf + a.offset + 4 * i, v);
      } else {
        for (int r = 0; r < a.world; ++r)
          if (r != root) st_f4(a.peer_bufs[r] + a.offset + 4 * i, v);
      }
    }
  }
  cross_gpu_barrier(a, 1, e, false);
  epoch_end(a, e);
}

template <int WIRE>
void launch_oneshot(const AllReduceArgs& a, int blocks, cudaStream_t stream) {
  if (a.mc_scratch != nullptr) oneshot_push_kernel<WIRE, true><<<blocks, kThreads, 0, stream>>>(a);
  else oneshot_push_kernel<WIRE, false><<<blocks, kThreads, 0, stream>>>(a);
}

}  // namespace

void launch_fused_allreduce(const AllReduceArgs& args, cudaStream_t stream) {
  if (args.world < 1 || args.world > kMaxRanks) throw std::runtime_error("fused_allreduce: world must be in [1, 8]");
  if (args.channel < 0 || args.channel >= kCommChannels) throw std::runtime_error("fused_allreduce: bad channel");
  const bool bf16 = args.wire == kWireBF16;
  if (args.wire != kWireF32 && !bf16) throw std::runtime_error("fused_allreduce: unknown wire format");
  const int vec = bf16 ? 8 : 4;
  const bool have_scratch = args.peer_scratch[args.rank] != nullptr;
  if (bf16 && !have_scratch) throw std::runtime_error("fused_allreduce: the bf16 wire format needs the scratch region");
  if (args.offset % vec != 0 || args.numel % vec != 0)
    throw std::runtime_error("fused_allreduce: bucket must be 16-byte (fp32 wire) / 32-byte (bf16 wire) aligned");
  int algo = args.algo;
  const bool twoshot_ok = args.numel % ((long long)vec * args.world) == 0;
  if (algo == kAlgoAuto) {
    if (have_scratch && args.numel <= kOneShotMaxElems) algo = kAlgoOneShot;
    else {
      // in-switch reduction pays from 4 GPUs up; between 2 GPUs every element would cross the switch twice for nothing
      // (measured at N = 2, 64 MB: peer loads 144 us, multimem 191 us, NCCL 164 us)
      const bool mc = bf16 ? args.mc_scratch != nullptr : args.mc_buf != nullptr;
      algo = (mc && args.world > 2) ? kAlgoMultimem : kAlgoTwoShot;
    }
  }
  if (algo == kAlgoOneShot && (!have_scratch || args.numel > kOneShotMaxElems))
    throw std::runtime_error("fused_allreduce: one-shot needs the scratch region and <= 64 Ki elements");
  if (algo != kAlgoOneShot && !twoshot_ok)
    throw std::runtime_error("fused_allreduce: two-shot buckets must be a multiple of (16 B x world) (plan_buckets pads them)");
  if (algo == kAlgoMultimem && (bf16 ? args.mc_scratch == nullptr : args.mc_buf == nullptr))
    throw std::runtime_error("fused_allreduce: no multicast mapping");
  AllReduceArgs a = args;
  if (a.timeout_ms <= 0) a.timeout_ms = 30000;
  if (algo == kAlgoOneShot) {
    // exactly one vector per thread: block b always owns bytes [4 KiB x b, 4 KiB x (b+1)) of every push slot, whatever
    // the bucket size, so the per-(channel, block) epoch parity is a valid double-buffer index for that range
    long long want = (args.numel / vec + kThreads - 1) / kThreads;
    const int blocks = (int)(want < 1 ? 1 : want);
    if (bf16) launch_oneshot<kWireBF16>(a, blocks, stream);
    else launch_oneshot<kWireF32>(a, blocks, stream);
  } else {
    const long long vec_per_rank = args.numel / args.world / vec;
    long long want = (vec_per_rank + kThreads - 1) / kThreads;
    // 64 CTAs of 256 threads keep ~2 MB in flight; from 1 MB per bucket up 128 CTAs pay (N = 2, 64 MB: 126 vs 168 us)
    int blocks = args.blocks > 0 ? args.blocks : (args.numel >= (1 << 18) ? 128 : 64);
    if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
    if (want < blocks) blocks = (int)(want > 0 ? want : 1);
    const bool mc = algo == kAlgoMultimem;
    if (bf16) {
      if (mc) twoshot_kernel<kWireBF16, true><<<blocks, kThreads, 0, stream>>>(a);
      else twoshot_kernel<kWireBF16, false><<<blocks, kThreads, 0, stream>>>(a);
    } else {
      if (mc) twoshot_kernel<kWireF32, true><<<blocks, kThreads, 0, stream>>>(a);
      else twoshot_kernel<kWireF32, false><<<blocks, kThreads, 0, stream>>>(a);
    }
  }
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_symm_broadcast(const AllReduceArgs& args, int root, cudaStream_t stream) {
  if (args.numel % 4 != 0 || args.offset % 4 != 0) throw std::runtime_error("symm_broadcast: region must be 16-byte aligned");
  if (args.channel < 0 || args.channel >= kCommChannels) throw std::runtime_error("symm_broadcast: bad channel");
  long long want = ((args.numel >> 2) + kThreads - 1) / kThreads;
  int blocks = (int)(want < 16 ? (want > 0 ? want : 1) : 16);
  AllReduceArgs a = args;
  if (a.timeout_ms <= 0) a.timeout_ms = 30000;
  if (a.mc_buf != nullptr) symm_broadcast_kernel<true><<<blocks, kThreads, 0, stream>>>(a, root);
  else symm_broadcast_kernel<false><<<blocks, kThreads, 0, stream>>>(a, root);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
