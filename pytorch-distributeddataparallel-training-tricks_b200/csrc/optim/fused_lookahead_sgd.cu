// Single-launch Lookahead(k, alpha) + SGD(momentum, nesterov) for sm_100a.
//
// Replaces, per optimizer step, ~100-150 launches of the reference stack (torch.optim.SGD's
// per-tensor / foreach kernels, reference utils/train.py:41, plus the four ATen ops per tensor
// of utils/lookahead.py:19-27 every k-th step) and the per-call fp32->fp16
// weight casts Apex O1 inserted before every conv/linear (N9): the bf16 shadow copy that the
// tcgen05 kernels read is written by this same kernel.
//
// All control state is device resident: learning rate and friends in ``hyper``, the step and
// Lookahead counters in ``state``, the amp overflow flag in ``found_inf``.  The kernel therefore
// has no data-dependent launch parameters and can sit inside a captured CUDA graph; an overflow
// turns the launch into a no-op that also leaves the Lookahead counter untouched (Apex skips the
// whole wrapped ``step``).
//
// Semantics (bit-compatible with torch.optim.SGD + reference Lookahead in fp32):
//   g   = grad + wd * p
//   buf = g                      (first step)       |  buf = mu * buf + (1 - damp) * g
//   d   = nesterov ? g + mu * buf : buf
//   p  -= lr * d
//   if la_counter == 0:  slow = (slow_ready ? slow + alpha * (p - slow) : p);  p = slow
//   la_counter = (la_counter + 1) % k
//
// Memory traffic per element: read p, g, buf (+ slow on sync steps), write p, buf (+ slow, +2 B
// bf16).  The kernel is a pure HBM stream: 128-bit accesses, 4 independent vectors in flight per
// thread, grid = 2 CTAs per SM x grid-stride.
#include "../include/ddpt_optim.h"
#include "../include/ddpt_common.cuh"

namespace ddpt {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// 16-bit weight shadow: bf16 (O1) or fp16 (O1-fp16)
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* dst, float4 v, int fp16 = 0) {
  uint2 packed;
  if (fp16) { packed.x = pack_f16x2(v.x, v.y); packed.y = pack_f16x2(v.z, v.w); }
  else { packed.x = pack_bf16x2(v.x, v.y); packed.y = pack_bf16x2(v.z, v.w); }
  *reinterpret_cast<uint2*>(dst) = packed;
}

__global__ void __launch_bounds__(kThreads)
fused_lookahead_sgd_kernel(FusedSgdArgs a) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  if (a.found_inf != nullptr && *a.found_inf != 0.0f) {
    if (blockIdx.x == 0 && threadIdx.x == 0) a.state[kSkippedSteps] += 1;
    return;  // overflow: skip the whole step, counters included (uniform across the grid)
  }
  const float lr = a.hyper[kLr], mu = a.hyper[kMomentum], wd = a.hyper[kWeightDecay];
  const float alpha = a.hyper[kAlpha], one_minus_damp = 1.0f - a.hyper[kDampening];
  const bool first_step = a.state[kStepCount] == 0;
  const bool lookahead = a.slow != nullptr;
  const bool sync = lookahead && a.state[kLaCounter] == 0;
  const bool slow_ready = lookahead && a.state[kSlowReady] != 0;
  const bool nesterov = a.nesterov != 0;

  const long long nvec = a.numel >> 2;
  float4* __restrict__ p4 = reinterpret_cast<float4*>(a.params);
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(a.grads);
  float4* __restrict__ m4 = reinterpret_cast<float4*>(a.momentum);
  float4* __restrict__ s4 = reinterpret_cast<float4*>(a.slow);
  float4* __restrict__ z4 = reinterpret_cast<float4*>(a.grads_to_zero);

  const long long stride = (long long)gridDim.x * kThreads;
  for (long long base = (long long)blockIdx.x * kThreads + threadIdx.x; base < nvec; base += stride * kUnroll) {
    float4 p[kUnroll], g[kUnroll], m[kUnroll], s[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = base + u * stride;
      if (i < nvec) {
        p[u] = p4[i];
        g[u] = g4[i];
        if (!first_step) m[u] = m4[i];
        if (sync && slow_ready) s[u] = s4[i];
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = base + u * stride;
      if (i >= nvec) continue;
      float* pp = reinterpret_cast<float*>(&p[u]);
      float* gg = reinterpret_cast<float*>(&g[u]);
      float* mm = reinterpret_cast<float*>(&m[u]);
      float* ss = reinterpret_cast<float*>(&s[u]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float grad = gg[c];
        if (wd != 0.0f) grad = fmaf(wd, pp[c], grad);
        float buf = first_step ? grad : fmaf(mu, mm[c], one_minus_damp * grad);
        float d = nesterov ? fmaf(mu, buf, grad) : buf;
        float w = fmaf(-lr, d, pp[c]);
        if (sync) {
          float sl = slow_ready ? fmaf(alpha, w - ss[c], ss[c]) : w;
          ss[c] = sl;
          w = sl;
        }
        mm[c] = buf;
        pp[c] = w;
      }
      p4[i] = p[u];
      m4[i] = m[u];
      if (sync) s4[i] = s[u];
      if (a.shadow != nullptr) st_bf16x4(a.shadow + 4 * i, p[u], a.shadow_fp16);
      if (z4 != nullptr) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // The last CTA to finish advances the counters (every CTA has read them by then).
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(&a.state[kTicket], 1);
    if (ticket == (int)gridDim.x - 1) {
      a.state[kTicket] = 0;
      a.state[kStepCount] += 1;
      if (lookahead) {
        if (sync) a.state[kSlowReady] = 1;
        int c = a.state[kLaCounter] + 1;
        a.state[kLaCounter] = (c >= a.state[kLaK]) ? 0 : c;
      }
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(kThreads)
cast_bf16_kernel(const float4* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long nvec, int fp16) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (long long)gridDim.x * kThreads)
    st_bf16x4(dst + 4 * i, src[i], fp16);
}

}  // namespace

void launch_fused_lookahead_sgd(const FusedSgdArgs& args, cudaStream_t stream) {
  if (args.numel <= 0) return;
  const long long nvec = args.numel >> 2;
  long long want = (nvec + (long long)kThreads * kUnroll - 1) / ((long long)kThreads * kUnroll);
  const int cap = sm_count() * 4;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  launch_pdl(fused_lookahead_sgd_kernel, dim3(grid), dim3(kThreads), 0, stream, args);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_cast_bf16(const float* src, __nv_bfloat16* dst, long long numel, cudaStream_t stream, int fp16) {
  const long long nvec = numel >> 2;
  if (nvec <= 0) return;
  long long want = (nvec + kThreads - 1) / kThreads;
  const int cap = sm_count() * 8;
  launch_pdl(cast_bf16_kernel, dim3((int)(want < cap ? want : cap)), dim3(kThreads), 0, stream, 
      reinterpret_cast<const float4*>(src), dst, nvec, fp16);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
