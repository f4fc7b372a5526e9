// Bandwidth-bound half of the fused Toy_Net step on sm_100a: first convolution (K = 9), BatchNorm
// statistics / normalise / ReLU / max-pool forward and backward, and the classifier head.
//
// Replaces, per training step, the cuDNN BatchNorm fwd/bwd, ATen relu_/max_pool2d(+backward),
// flatten, log_softmax/nll_loss/argmax/eq/sum kernels of the reference stack (SURVEY.md N11, N12,
// N14; Toy_Net ops K1, K2, K4, K5, K7, K9, K10, K12-K14) with a handful of fused passes:
//
//   forward : conv epilogue -> stat partials -> bn_finalize -> bn_relu_pool   (1 read + <=1 write)
//   backward: bn_bwd_reduce (1 read of y and g) -> bn_bwd_finalize -> bn_bwd_apply (1 read, 1 write)
//
// All activations are NHWC bf16; every thread moves 16-byte vectors (8 channels) and keeps to a
// fixed channel group so per-channel reductions stay in registers until one block-level fold.
#include "../include/ddpt_toynet.h"
#include "../include/ddpt_common.cuh"
#include <cuda_fp8.h>
#include <cstdlib>
#include <stdexcept>

// ---- activation precision -------------------------------------------------------------------------------------------
// This translation unit is compiled twice: as is for bf16 activations (the default "O1"), and through
// toynet_elementwise_f16.cu (which defines DDPT_ACT_F16 and includes this file) for fp16 activations -- Apex O1's actual
// precision in the reference (utils/train.py:58).  Every kernel is written against act_t / act2_t and the six helpers
// below; the public launchers are C++ overloads on the activation pointer type, so both builds link into one library.
#ifdef DDPT_ACT_F16
#include <cuda_fp16.h>
#endif

namespace ddpt {
namespace {

#ifdef DDPT_ACT_F16
using act_t = __half;
using act2_t = __half2;
__device__ __forceinline__ float2 act2_to_float2(act2_t v) { return __half22float2(v); }
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ act_t to_act(float f) { return __float2half_rn(f); }
__device__ __forceinline__ act_t to_act_rd(float f) { return __float2half_rd(f); }
__device__ __forceinline__ float from_act(act_t v) { return __half2float(v); }
__device__ __forceinline__ float bf_lo(uint32_t x) { return __half2float(__ushort_as_half((unsigned short)(x & 0xffffu))); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __half2float(__ushort_as_half((unsigned short)(x >> 16))); }
#else
using act_t = __nv_bfloat16;
using act2_t = __nv_bfloat162;
__device__ __forceinline__ float2 act2_to_float2(act2_t v) { return __bfloat1622float2(v); }
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) { return pack_bf16x2(lo, hi); }
__device__ __forceinline__ act_t to_act(float f) { return __float2bfloat16(f); }
__device__ __forceinline__ act_t to_act_rd(float f) { return __float2bfloat16_rd(f); }
__device__ __forceinline__ float from_act(act_t v) { return __bfloat162float(v); }
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }          // low / high element of a packed pair
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }
#endif

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const act2_t* p = reinterpret_cast<const act2_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = act2_to_float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_act2(f[0], f[1]);
  v.y = pack_act2(f[2], f[3]);
  v.z = pack_act2(f[4], f[5]);
  v.w = pack_act2(f[6], f[7]);
  return v;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// MXFP8 side output of the activation passes ("O1-fp8"): shared UE8M0 exponent of a 32-element block from its amax.
__device__ __forceinline__ int mx_exponent(float amax) {
  int e = 127;
  if (amax > 0.f && is_finite_f32(amax)) {
    int ex;
    const float m = frexpf(amax * (1.0f / 448.0f), &ex);
    e = 127 + ex - (m == 0.5f ? 1 : 0);
    e = max(1, min(254, e));
  }
  return e;
}
__device__ __forceinline__ uint32_t e4m3x4(float a, float b, float c, float d) {
  return (uint32_t)__nv_cvt_float_to_fp8(a, __NV_SATFINITE, __NV_E4M3) |
         ((uint32_t)__nv_cvt_float_to_fp8(b, __NV_SATFINITE, __NV_E4M3) << 8) |
         ((uint32_t)__nv_cvt_float_to_fp8(c, __NV_SATFINITE, __NV_E4M3) << 16) |
         ((uint32_t)__nv_cvt_float_to_fp8(d, __NV_SATFINITE, __NV_E4M3) << 24);
}

int persistent_grid(long long work_items, int per_block, int blocks_per_sm) {
  long long want = (work_items + per_block - 1) / per_block;
  long long cap = (long long)sm_count() * blocks_per_sm;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ================================================================================================
// conv.0 : direct 3x3, C_in = 1, C_out = 64
// ================================================================================================
constexpr int kImg = 28, kOut0 = 26, kPix0 = kOut0 * kOut0, kC0 = 64;

__device__ __forceinline__ void load_image(const void* x, int x_is_u8, int img, float* s_img) {
  if (x_is_u8) {
    const uint8_t* src = static_cast<const uint8_t*>(x) + (size_t)img * kImg * kImg;
    for (int i = threadIdx.x; i < kImg * kImg / 4; i += kThreads) {
      uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
      s_img[4 * i] = v.x * (1.0f / 255.0f);
      s_img[4 * i + 1] = v.y * (1.0f / 255.0f);
      s_img[4 * i + 2] = v.z * (1.0f / 255.0f);
      s_img[4 * i + 3] = v.w * (1.0f / 255.0f);
    }
  } else {
    const float* src = static_cast<const float*>(x) + (size_t)img * kImg * kImg;
    for (int i = threadIdx.x; i < kImg * kImg; i += kThreads) s_img[i] = src[i];
  }
}

__global__ void __launch_bounds__(kThreads)
conv0_fwd_kernel(const void* __restrict__ x, int x_is_u8, const float* __restrict__ weight,
                 const float* __restrict__ bias, act_t* __restrict__ y, float* __restrict__ partials,
                 double* __restrict__ accum, int batch) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  __shared__ float s_img[kImg * kImg];
  __shared__ float s_red[kThreads / 32][2][kC0];            // one slot per warp: the block fold runs in a fixed order
  const int cg = threadIdx.x & 7, prow = threadIdx.x >> 3;
  float wr[8][9], br[8], acc_s[8], acc_q[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
    acc_s[c] = acc_q[c] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = weight[(cg * 8 + c) * 9 + t];
  }
  for (int img = blockIdx.x; img < batch; img += gridDim.x) {
    __syncthreads();
    load_image(x, x_is_u8, img, s_img);
    __syncthreads();
    for (int p = prow; p < kPix0; p += kThreads / 8) {
      const int oh = p / kOut0, ow = p - oh * kOut0;
      float in[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) in[r * 3 + q] = s_img[(oh + r) * kImg + ow + q];
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = br[c];
#pragma unroll
        for (int t = 0; t < 9; ++t) v = fmaf(wr[c][t], in[t], v);
        o[c] = v;
        acc_s[c] += v;
        acc_q[c] = fmaf(v, v, acc_q[c]);
      }
      *reinterpret_cast<uint4*>(y + ((size_t)img * kPix0 + p) * kC0 + cg * 8) = pack8(o);
    }
  }
  if (partials == nullptr && accum == nullptr) return;
  // fold the 32 threads sharing a channel group: 4 lanes per warp (xor 8, 16), then the 8 warps in warp order
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float s = acc_s[c], q = acc_q[c];
    s += __shfl_xor_sync(0xffffffffu, s, 8);  q += __shfl_xor_sync(0xffffffffu, q, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16); q += __shfl_xor_sync(0xffffffffu, q, 16);
    if ((threadIdx.x & 31) < 8) {
      s_red[warp][0][cg * 8 + c] = s;
      s_red[warp][1][cg * 8 + c] = q;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kC0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot += (&s_red[w][0][0])[threadIdx.x];
    if (accum != nullptr) atomicAdd(accum + threadIdx.x, (double)tot);
    else partials[(size_t)blockIdx.x * 2 * kC0 + threadIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(kThreads)
conv0_wgrad_kernel(const void* __restrict__ x, int x_is_u8, const act_t* __restrict__ dy,
                   float* __restrict__ dw, float* __restrict__ db, float* __restrict__ partials, int batch,
                   const Conv0BnFuse bn) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  __shared__ float s_img[kImg * kImg];
  __shared__ float s_acc[kThreads / 32][kC0 * 10];         // one slot per warp: ordered block fold
  const int cg = threadIdx.x & 7, prow = threadIdx.x >> 3;
  // BatchNorm-backward "apply" of block 0 folded in: conv.0 has no data gradient, so dy0 is read exactly once -- here.
  // Instead of a pass that reads y0 and g and writes dy0 (88 MB) for this kernel to read back, compute
  // dy = c1 * dz + k0 - k1 * y (dz = g where relu(bn(y)) > 0) from y0 and g directly, rounded to bf16 like the stored one.
  const bool fused = bn.y != nullptr;
  float f_sc[8], f_sh[8], f_c1[8], f_k0[8], f_k1[8];
  if (fused) {
    float mu[8], is[8], c2[8], c3[8];
    load8f(bn.scale + cg * 8, f_sc); load8f(bn.shift + cg * 8, f_sh);
    load8f(bn.mean + cg * 8, mu); load8f(bn.invstd + cg * 8, is);
    load8f(bn.bcoef + cg * 8, f_c1); load8f(bn.bcoef + kC0 + cg * 8, c2); load8f(bn.bcoef + 2 * kC0 + cg * 8, c3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f_k1[j] = f_c1[j] * is[j] * c3[j];
      f_k0[j] = fmaf(f_k1[j], mu[j], -f_c1[j] * c2[j]);
    }
  }
  float acc[8][10];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[c][t] = 0.f;
  for (int img = blockIdx.x; img < batch; img += gridDim.x) {
    __syncthreads();
    load_image(x, x_is_u8, img, s_img);
    __syncthreads();
    constexpr int kStep = kThreads / 8;                      // 32 pixels per sweep of the block
    for (int p0 = prow; p0 < kPix0; p0 += 4 * kStep) {
      uint4 raw[4], rawy[4];                                 // four gradient vectors in flight per thread
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kStep;
        if (p < kPix0) {
          const size_t off = ((size_t)img * kPix0 + p) * kC0 + cg * 8;
          if (fused) { raw[u] = ldg16(static_cast<const act_t*>(bn.g) + off); rawy[u] = ldg16(static_cast<const act_t*>(bn.y) + off); }
          else raw[u] = ldg16(dy + off);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kStep;
        if (p >= kPix0) continue;
        const int oh = p / kOut0, ow = p - oh * kOut0;
        float g[8];
        unpack8(raw[u], g);
        if (fused) {
          float yv[8];
          unpack8(rawy[u], yv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float dz = fmaf(yv[j], f_sc[j], f_sh[j]) > 0.f ? g[j] : 0.f;
            g[j] = from_act(to_act(fmaf(f_c1[j], dz, fmaf(-f_k1[j], yv[j], f_k0[j]))));
          }
        }
        float in[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) in[r * 3 + q] = s_img[(oh + r) * kImg + ow + q];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[c][t] = fmaf(g[c], in[t], acc[c][t]);
          acc[c][9] += g[c];
        }
      }
    }
  }
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      float v = acc[c][t];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) s_acc[warp][(cg * 8 + c) * 10 + t] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < kC0 * 10; i += kThreads) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) tot += s_acc[w][i];
    const int ch = i / 10, t = i - ch * 10;
    if (partials != nullptr) {
      // deterministic mode: one row per block ([576 dw | 64 db]), folded in block order by reduce_rows_kernel
      partials[(size_t)blockIdx.x * (kC0 * 10) + (t < 9 ? ch * 9 + t : kC0 * 9 + ch)] = tot;
    } else if (t < 9) {
      atomicAdd(dw + ch * 9 + t, tot);
    } else {
      atomicAdd(db + ch, tot);
    }
  }
}

// Same contract, 4 channels per thread and 512 threads per CTA: 40 accumulators instead of 80 (~100 registers), so the
// SM holds 16 warps instead of 8.  The 8-channel kernel is issue-bound with two warps per scheduler at ~4.9 cycles per
// instruction (profiles/ncu_r2_conv0_wgrad.txt); twice the warps hide the FFMA / LDS latencies, at the price of twice the
// patch loads per FFMA.  DDPT_CONV0_WGRAD_CH=4 selects it.
constexpr int kThreads4 = 512;

__device__ __forceinline__ void unpack4(const uint2& v, float (&f)[4]) {
  const act2_t* p = reinterpret_cast<const act2_t*>(&v);
  float2 a = act2_to_float2(p[0]), b = act2_to_float2(p[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__global__ void __launch_bounds__(kThreads4, 1)
conv0_wgrad4_kernel(const void* __restrict__ x, int x_is_u8, const act_t* __restrict__ dy,
                    float* __restrict__ dw, float* __restrict__ db, float* __restrict__ partials, int batch,
                    const Conv0BnFuse bn) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_img[kImg * kImg];
  __shared__ float s_acc[kThreads4 / 32][kC0 * 10];        // one slot per warp: ordered block fold
  const int cg = threadIdx.x & 15, prow = threadIdx.x >> 4;   // 16 channel groups of 4, 32 pixel rows
  const bool fused = bn.y != nullptr;
  float f_sc[4], f_sh[4], f_c1[4], f_k0[4], f_k1[4];
  if (fused) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cg * 4 + j;
      const float c1 = __ldg(bn.bcoef + c), c2 = __ldg(bn.bcoef + kC0 + c), c3 = __ldg(bn.bcoef + 2 * kC0 + c);
      f_sc[j] = __ldg(bn.scale + c); f_sh[j] = __ldg(bn.shift + c);
      f_c1[j] = c1;
      f_k1[j] = c1 * __ldg(bn.invstd + c) * c3;
      f_k0[j] = fmaf(f_k1[j], __ldg(bn.mean + c), -c1 * c2);
    }
  }
  float acc[4][10];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[c][t] = 0.f;
  for (int img = blockIdx.x; img < batch; img += gridDim.x) {
    __syncthreads();
    if (x_is_u8) {
      const uint8_t* src = static_cast<const uint8_t*>(x) + (size_t)img * kImg * kImg;
      for (int i = threadIdx.x; i < kImg * kImg / 4; i += kThreads4) {
        uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
        s_img[4 * i] = v.x * (1.0f / 255.0f); s_img[4 * i + 1] = v.y * (1.0f / 255.0f);
        s_img[4 * i + 2] = v.z * (1.0f / 255.0f); s_img[4 * i + 3] = v.w * (1.0f / 255.0f);
      }
    } else {
      const float* src = static_cast<const float*>(x) + (size_t)img * kImg * kImg;
      for (int i = threadIdx.x; i < kImg * kImg; i += kThreads4) s_img[i] = src[i];
    }
    __syncthreads();
    constexpr int kStep = kThreads4 / 16;                    // 32 pixels per sweep of the block
    for (int p0 = prow; p0 < kPix0; p0 += 4 * kStep) {
      uint2 raw[4], rawy[4];                                 // four gradient vectors in flight per thread
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kStep;
        if (p < kPix0) {
          const size_t off = ((size_t)img * kPix0 + p) * kC0 + cg * 4;
          if (fused) {
            raw[u] = __ldg(reinterpret_cast<const uint2*>(static_cast<const act_t*>(bn.g) + off));
            rawy[u] = __ldg(reinterpret_cast<const uint2*>(static_cast<const act_t*>(bn.y) + off));
          } else {
            raw[u] = __ldg(reinterpret_cast<const uint2*>(dy + off));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kStep;
        if (p >= kPix0) continue;
        const int oh = p / kOut0, ow = p - oh * kOut0;
        float g[4];
        unpack4(raw[u], g);
        if (fused) {
          float yv[4];
          unpack4(rawy[u], yv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float dz = fmaf(yv[j], f_sc[j], f_sh[j]) > 0.f ? g[j] : 0.f;
            g[j] = from_act(to_act(fmaf(f_c1[j], dz, fmaf(-f_k1[j], yv[j], f_k0[j]))));
          }
        }
        float in[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) in[r * 3 + q] = s_img[(oh + r) * kImg + ow + q];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[c][t] = fmaf(g[c], in[t], acc[c][t]);
          acc[c][9] += g[c];
        }
      }
    }
  }
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      float v = acc[c][t];
      v += __shfl_xor_sync(0xffffffffu, v, 16);               // the two pixel rows of a warp that share a channel group
      if ((threadIdx.x & 31) < 16) s_acc[warp][(cg * 4 + c) * 10 + t] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < kC0 * 10; i += kThreads4) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads4 / 32; ++w) tot += s_acc[w][i];
    const int ch = i / 10, t = i - ch * 10;
    if (partials != nullptr) {
      partials[(size_t)blockIdx.x * (kC0 * 10) + (t < 9 ? ch * 9 + t : kC0 * 9 + ch)] = tot;
    } else if (t < 9) {
      atomicAdd(dw + ch * 9 + t, tot);
    } else {
      atomicAdd(db + ch, tot);
    }
  }
}

// out[i] (+)= sum over rows r (in order) of partials[r * ld + i]: the ordered cross-block fold of deterministic mode
__global__ void __launch_bounds__(kThreads)
reduce_rows_kernel(const float* __restrict__ partials, int rows, int ld, int n, float* __restrict__ out, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;             // fixed association: four interleaved chains, then a tree
  int r = 0;
  for (; r + 3 < rows; r += 4) {
    a0 += partials[(size_t)r * ld + i];
    a1 += partials[(size_t)(r + 1) * ld + i];
    a2 += partials[(size_t)(r + 2) * ld + i];
    a3 += partials[(size_t)(r + 3) * ld + i];
  }
  for (; r < rows; ++r) a0 += partials[(size_t)r * ld + i];
  const float tot = (a0 + a1) + (a2 + a3);
  out[i] = accumulate ? out[i] + tot : tot;
}

// ================================================================================================
// BatchNorm forward
// ================================================================================================
// One block per 32 channels; 8 row-groups of 32 lanes walk the partial rows (coalesced across channels).
constexpr int kFinGroups = 32;                       // row groups per block in the *_finalize kernels

__global__ void __launch_bounds__(32 * kFinGroups)
bn_finalize_kernel(const float* __restrict__ partials, int num_partials, int channels, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                   BnCoeffs out, float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ nbt, const float* __restrict__ conv_bias) {
  // conv_bias (optional): the statistics are those of the convolution WITHOUT its bias -- the producing GEMM skipped the
  // add (a per-channel constant in front of a BatchNorm cancels in the normalised output and in every gradient; only
  // running_mean sees it).  mean / shift stay consistent with the stored bias-free activations.
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  __shared__ double s_sum[kFinGroups][32], s_sq[kFinGroups][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double sum = 0.0, sq = 0.0;
  if (c < channels) {
    // 4 row groups per trip with independent accumulators: the loads of a trip are all in flight together
    double s4[4] = {0.0, 0.0, 0.0, 0.0}, q4[4] = {0.0, 0.0, 0.0, 0.0};
    int p = grp;
    for (; p + 3 * kFinGroups < num_partials; p += 4 * kFinGroups) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = partials[((size_t)(p + u * kFinGroups) * 2) * channels + c];
        b[u] = partials[((size_t)(p + u * kFinGroups) * 2 + 1) * channels + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s4[u] += (double)a[u]; q4[u] += (double)b[u]; }
    }
    for (; p < num_partials; p += kFinGroups) {
      s4[0] += (double)partials[((size_t)p * 2) * channels + c];
      q4[0] += (double)partials[((size_t)p * 2 + 1) * channels + c];
    }
    sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
  }
  s_sum[grp][lane] = sum;
  s_sq[grp][lane] = sq;
  __syncthreads();
  if (grp == 0 && c < channels) {
    for (int g = 1; g < kFinGroups; ++g) { sum += s_sum[g][lane]; sq += s_sq[g][lane]; }
    const double mean = sum / count;
    double var = sq / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    out.scale[c] = sc;
    out.shift[c] = beta[c] - (float)mean * sc;
    out.mean[c] = (float)mean;
    out.invstd[c] = invstd;
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      const float batch_mean = conv_bias != nullptr ? (float)mean + conv_bias[c] : (float)mean;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * batch_mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
  if (nbt != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      int channels, float* __restrict__ scale, float* __restrict__ shift,
                                      const float* __restrict__ conv_bias) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  const float sc = gamma[c] * rsqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - (conv_bias != nullptr ? rm[c] - conv_bias[c] : rm[c]) * sc;   // bias-free activations (see bn_finalize)
}

// Fused statistics (BnFusedFwd): the first C/8 threads of every CTA turn the fp64 accumulators into scale / shift for
// 8 channels each (shared memory, read by everyone after the barrier); CTA 0 also publishes scale / shift / mean /
// invstd for the backward pass and updates the running statistics (momentum, unbiased variance) and the step counter.
__device__ __forceinline__ void fused_fwd_coeffs(const BnFusedFwd& st, int c, float* s_sc, float* s_sh) {
  const int cv = c >> 3;
  if ((int)threadIdx.x < cv) {
    const bool publish = blockIdx.x == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = threadIdx.x * 8 + j;
      // fp64 only where cancellation matters (E[x^2] - mean^2); one multiply and one fma, no division, no sqrt
      const double mean = st.accum[ch] * st.inv_count;
      const float var = fmaxf((float)fma(st.accum[c + ch], st.inv_count, -mean * mean), 0.f);
      const float invstd = rsqrtf(var + st.eps);
      const float sc = st.gamma[ch] * invstd;
      const float sh = st.beta[ch] - (float)mean * sc;
      s_sc[ch] = sc;
      s_sh[ch] = sh;
      if (publish) {
        st.out.scale[ch] = sc;
        st.out.shift[ch] = sh;
        st.out.mean[ch] = (float)mean;
        st.out.invstd[ch] = invstd;
        if (st.running_mean != nullptr) {
          st.running_mean[ch] = (1.f - st.momentum) * st.running_mean[ch] + st.momentum * (float)mean;
          st.running_var[ch] = (1.f - st.momentum) * st.running_var[ch] + st.momentum * (var * st.unbias);
        }
      }
    }
    if (publish && threadIdx.x == 0 && st.num_batches_tracked != nullptr) *st.num_batches_tracked += 1;
  }
  __syncthreads();
}
__device__ __forceinline__ void load8s(const float* p, float (&f)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = p[j];
}

// a = relu(y*scale+shift) [+ 2x2 max-pool]; a thread owns one 8-channel group and walks output pixels; the raw
// loads of U pixels are issued before any arithmetic (these passes are latency-, not bandwidth-limited otherwise).
template <bool POOL>
__global__ void __launch_bounds__(kThreads)
bn_relu_pool_kernel(const act_t* __restrict__ y, const float* __restrict__ scale,
                    const float* __restrict__ shift, act_t* __restrict__ out, int n, int h, int w, int c,
                    uint8_t* __restrict__ q_out, uint8_t* __restrict__ q_scale, const BnFusedFwd st) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  __shared__ float s_sc[512], s_sh[512];
  constexpr int U = POOL ? 2 : 4;
  const int cv = c >> 3;
  const int v = threadIdx.x % cv, lane_row = threadIdx.x / cv, rows_per_blk = kThreads / cv;
  const int oh = POOL ? h >> 1 : h, ow = POOL ? w >> 1 : w;
  const unsigned total = (unsigned)n * oh * ow;
  const unsigned stride = gridDim.x * rows_per_blk;
  float sc[8], sh[8];
  if (st.accum != nullptr) {
    fused_fwd_coeffs(st, c, s_sc, s_sh);
    load8s(s_sc + v * 8, sc);
    load8s(s_sh + v * 8, sh);
  } else {
    load8f(scale + v * 8, sc);
    load8f(shift + v * 8, sh);
  }
  for (unsigned pix0 = blockIdx.x * rows_per_blk + lane_row; pix0 < total; pix0 += U * stride) {
    uint4 raw[U][POOL ? 4 : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned pix = pix0 + u * stride;
      if (pix >= total) continue;
      if (!POOL) {
        raw[u][0] = ldg16(y + (size_t)pix * c + v * 8);
      } else {
        const unsigned pw = pix % ow, ph = (pix / ow) % oh, img = pix / ((unsigned)ow * oh);
        const act_t* base = y + (((size_t)img * h + 2 * ph) * w + 2 * pw) * c + v * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[u][q] = ldg16(base + ((q >> 1) * w + (q & 1)) * c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned pix = pix0 + u * stride;
      if (pix >= total) continue;
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = 0.f;              // relu floor
#pragma unroll
      for (int q = 0; q < (POOL ? 4 : 1); ++q) {
        float f[8];
        unpack8(raw[u][q], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = fmaxf(r[k], fmaf(f[k], sc[k], sh[k]));
      }
      *reinterpret_cast<uint4*>(out + (size_t)pix * c + v * 8) = pack8(r);
      if (q_out != nullptr) {
        // block of 32 channels = this lane and its 3 neighbours (same pixel: C/8 is a multiple of 4)
        float amax = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) amax = fmaxf(amax, r[k]);          // r >= 0 after ReLU
        const unsigned grp_mask = __activemask();                       // the 4 lanes of a block share `pix`
        amax = fmaxf(amax, __shfl_xor_sync(grp_mask, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(grp_mask, amax, 2));
        const int e = mx_exponent(amax);
        const float inv = exp2f((float)(127 - e));
        uint2 pk;
        pk.x = e4m3x4(r[0] * inv, r[1] * inv, r[2] * inv, r[3] * inv);
        pk.y = e4m3x4(r[4] * inv, r[5] * inv, r[6] * inv, r[7] * inv);
        *reinterpret_cast<uint2*>(q_out + (size_t)pix * c + v * 8) = pk;
        if ((v & 3) == 0) q_scale[(size_t)pix * (c >> 5) + (v >> 2)] = (uint8_t)e;
      }
    }
  }
}

// Pooled + flattened in (c, ph, pw) order for the first Linear: one block per image, smem transpose.
__global__ void __launch_bounds__(kThreads)
bn_relu_pool_flatten_kernel(const act_t* __restrict__ y, const float* __restrict__ scale,
                            const float* __restrict__ shift, act_t* __restrict__ out, int h, int w, int c,
                            uint8_t* __restrict__ q_out, uint8_t* __restrict__ q_scale, const BnFusedFwd st) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  extern __shared__ act_t s_t[];                  // [c][ppix + 2] (padded)
  __shared__ float s_sc[512], s_sh[512];
  const bool fused = st.accum != nullptr;
  if (fused) fused_fwd_coeffs(st, c, s_sc, s_sh);
  const int oh = h >> 1, ow = w >> 1, ppix = oh * ow, cv = c >> 3, ld = ppix + 2;
  const long long img = blockIdx.x;
  for (int i = threadIdx.x; i < ppix * cv; i += kThreads) {
    const int v = i % cv, pp = i / cv;
    const int ph = pp / ow, pw = pp - ph * ow;
    float sc[8], sh[8], r[8];
    if (fused) { load8s(s_sc + v * 8, sc); load8s(s_sh + v * 8, sh); }
    else { load8f(scale + v * 8, sc); load8f(shift + v * 8, sh); }
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = 0.f;
    const act_t* base = y + ((img * h + 2 * ph) * w + 2 * pw) * c + v * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
      unpack8(ldg16(base + ((q >> 1) * w + (q & 1)) * c), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = fmaxf(r[k], fmaf(f[k], sc[k], sh[k]));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s_t[(v * 8 + k) * ld + pp] = to_act(r[k]);
  }
  __syncthreads();
  act_t* dst = out + img * (long long)c * ppix;
  for (int i = threadIdx.x; i < c * ppix; i += kThreads) dst[i] = s_t[(i / ppix) * ld + (i % ppix)];
  if (q_out != nullptr) {
    // MXFP8 copy of the flattened row: blocks of 32 consecutive outputs (the K dimension of the first Linear)
    const int blocks = c * ppix / 32;
    for (int b = threadIdx.x; b < blocks; b += kThreads) {
      float vals[32];
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int i = b * 32 + j;
        vals[j] = from_act(s_t[(i / ppix) * ld + (i % ppix)]);
        amax = fmaxf(amax, vals[j]);
      }
      const int e = mx_exponent(amax);
      const float inv = exp2f((float)(127 - e));
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = e4m3x4(vals[4 * j] * inv, vals[4 * j + 1] * inv, vals[4 * j + 2] * inv, vals[4 * j + 3] * inv);
      uint4* o = reinterpret_cast<uint4*>(q_out + img * (long long)c * ppix + b * 32);
      o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      q_scale[img * (long long)blocks + b] = (uint8_t)e;
    }
  }
}

// ================================================================================================
// BatchNorm backward
// ================================================================================================
// dz for one 8-channel vector of one (pooled) output position; also returns xhat at the routed position(s).
struct BwdCtx {
  const act_t* y; const act_t* g;
  const float* scale; const float* shift; const float* mean; const float* invstd;
  int n, h, w, c, flatten;
};

__device__ __forceinline__ void load_g(const BwdCtx& k, long long img, int ph, int pw, int oh, int ow, int v,
                                       float (&g)[8]) {
  if (!k.flatten) {
    unpack8(ldg16(k.g + ((img * oh + ph) * ow + pw) * k.c + v * 8), g);
  } else {
    const act_t* base = k.g + img * (long long)k.c * oh * ow + (long long)(v * 8) * oh * ow + ph * ow + pw;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = from_act(base[(long long)j * oh * ow]);
  }
}

// ---- packed-bf16 arithmetic for the pooled layers ------------------------------------------------------------
// The 2x2 window arg-max and the ReLU gate are decided on the raw bf16 conv outputs, two channels per instruction:
// z = y*sc + sh is monotone in y, so with y' = y ^ sign(sc) (sign bit flipped where sc < 0) the winner is max(y')
// and z > 0  <=>  y' > thr' with thr' = floor_bf16(-sh/|sc|) (exact for bf16 y; -inf / +inf when sc == 0).
// ncu on the fp32 formulation: 28.5 M warp instructions, IPC 1.5, DRAM 34 % busy (profiles/ncu_bnreduce.txt) --
// the pooled passes were issue-bound, not bandwidth-bound; this form needs ~3x fewer instructions and half the
// registers per window in flight.
struct PoolConsts { uint32_t sgn[4], thr[4]; };

__device__ __forceinline__ PoolConsts pool_consts(const float (&sc)[8], const float (&sh)[8]) {
  PoolConsts pc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t sg = 0, th = 0;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const float a = sc[2 * i + hlf], b = sh[2 * i + hlf];
      const float t = a == 0.f ? (b > 0.f ? -INFINITY : INFINITY) : -b / fabsf(a);
      const act_t tb = to_act_rd(t);
      th |= (uint32_t)(*reinterpret_cast<const unsigned short*>(&tb)) << (16 * hlf);
      sg |= (a < 0.f ? 0x8000u : 0u) << (16 * hlf);
    }
    pc.sgn[i] = sg;
    pc.thr[i] = th;
  }
  return pc;
}
__device__ __forceinline__ act2_t as_bf162(uint32_t x) { return *reinterpret_cast<act2_t*>(&x); }
__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b) {
  const act2_t r = __hmax2(as_bf162(a), as_bf162(b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t comp(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// pooled gradient of one output position as four bf16x2 words (NHWC rows, or the flattened [c][oh][ow] layout
// the first Linear's data gradient is produced in)
__device__ __forceinline__ uint4 load_g_packed(const BwdCtx& k, int img, int ph, int pw, int oh, int ow, int v) {
  if (!k.flatten) return ldg16(k.g + (((size_t)img * oh + ph) * ow + pw) * k.c + v * 8);
  const unsigned short* base = reinterpret_cast<const unsigned short*>(k.g) + (size_t)img * k.c * oh * ow +
                               (size_t)(v * 8) * oh * ow + ph * ow + pw;
  const int plane = oh * ow;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = (uint32_t)__ldg(base + (2 * i) * plane) | ((uint32_t)__ldg(base + (2 * i + 1) * plane) << 16);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool POOL>
__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_reduce_kernel(BwdCtx k, float* __restrict__ partials, double* __restrict__ accum) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  extern __shared__ float s_part[];                       // [rows_per_blk][2][c]: ordered block fold (no float atomics)
  const int cv = k.c >> 3;
  const int v = threadIdx.x % cv, lane_row = threadIdx.x / cv, rows_per_blk = kThreads / cv;
  const int oh = POOL ? k.h >> 1 : k.h, ow = POOL ? k.w >> 1 : k.w;
  const long long total_pix = (long long)k.n * oh * ow;
  // row 1 of the partials holds sum(dz * y); bn_bwd_finalize turns it into sum(dz * xhat) = invstd * (sum(dz*y) - mean*sum(dz))
  float sc[8], sh[8], s_dz[8], s_dzx[8];
  load8f(k.scale + v * 8, sc); load8f(k.shift + v * 8, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) s_dz[j] = s_dzx[j] = 0.f;
  const long long stride = (long long)gridDim.x * rows_per_blk;
  if constexpr (POOL) {
    // output positions are walked with incremental 32-bit (pw, ph, img) updates -- no 64-bit divisions in the loop --
    // and the raw loads of U windows stay in flight at once
    constexpr int U = 3;
    const PoolConsts pc = pool_consts(sc, sh);
    const unsigned total32 = (unsigned)total_pix, stride32 = (unsigned)stride;
    const unsigned first = blockIdx.x * rows_per_blk + lane_row;
    const unsigned plane = (unsigned)ow * oh;
    int pw[U], ph[U], img[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = first + u * stride32;
      pw[u] = (int)(p % ow); ph[u] = (int)((p / ow) % oh); img[u] = (int)(p / plane);
    }
    const unsigned step = U * stride32;
    const int d_pw = (int)(step % ow), d_ph = (int)((step / ow) % oh), d_img = (int)(step / plane);
    for (unsigned pix = first; pix < total32; pix += step) {
      uint4 yq[U][4], gq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pix + u * stride32 < total32) {
          const act_t* base = k.y + (((size_t)img[u] * k.h + 2 * ph[u]) * k.w + 2 * pw[u]) * k.c + v * 8;
#pragma unroll
          for (int q = 0; q < 4; ++q) yq[u][q] = ldg16(base + ((q >> 1) * k.w + (q & 1)) * k.c);
          gq[u] = load_g_packed(k, img[u], ph[u], pw[u], oh, ow, v);
        } else {
          gq[u] = make_uint4(0, 0, 0, 0);                  // no gradient -> contributes nothing
#pragma unroll
          for (int q = 0; q < 4; ++q) yq[u][q] = make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t sg = pc.sgn[i];
          const uint32_t best = bmax2(bmax2(comp(yq[u][0], i) ^ sg, comp(yq[u][1], i) ^ sg),
                                      bmax2(comp(yq[u][2], i) ^ sg, comp(yq[u][3], i) ^ sg));
          const uint32_t dz = comp(gq[u], i) & __hgt2_mask(as_bf162(best), as_bf162(pc.thr[i]));
          const uint32_t by = best ^ sg;                   // winner's raw conv output
          s_dz[2 * i] += bf_lo(dz);
          s_dz[2 * i + 1] += bf_hi(dz);
          s_dzx[2 * i] = fmaf(bf_lo(dz), bf_lo(by), s_dzx[2 * i]);
          s_dzx[2 * i + 1] = fmaf(bf_hi(dz), bf_hi(by), s_dzx[2 * i + 1]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pw[u] += d_pw; if (pw[u] >= ow) { pw[u] -= ow; ++ph[u]; }
        ph[u] += d_ph; if (ph[u] >= oh) { ph[u] -= oh; ++img[u]; }
        img[u] += d_img;
      }
    }
  } else if (!k.flatten) {
    constexpr int U = 4;                                  // independent 16-byte loads in flight per thread: 2*U
    for (long long pix0 = (long long)blockIdx.x * rows_per_blk + lane_row; pix0 < total_pix; pix0 += stride * U) {
      uint4 yv[U], gv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pix = pix0 + u * stride;
        if (pix < total_pix) {
          yv[u] = ldg16(k.y + pix * k.c + v * 8);
          gv[u] = ldg16(k.g + pix * k.c + v * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pix0 + u * stride >= total_pix) continue;
        float f[8], g[8];
        unpack8(yv[u], f);
        unpack8(gv[u], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(f[j], sc[j], sh[j]);
          const float dz = z > 0.f ? g[j] : 0.f;
          s_dz[j] += dz;
          s_dzx[j] = fmaf(dz, f[j], s_dzx[j]);
        }
      }
    }
  } else {
    // un-pooled activation whose gradient arrives flattened ([c][h][w] per image): small layers only
    const int plane = oh * ow;
    for (long long pix = (long long)blockIdx.x * rows_per_blk + lane_row; pix < total_pix; pix += stride) {
      const int img = (int)(pix / plane), rem = (int)(pix % plane);
      float f[8], g[8];
      unpack8(ldg16(k.y + pix * k.c + v * 8), f);
      load_g(k, img, rem / ow, rem % ow, oh, ow, v, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(f[j], sc[j], sh[j]);
        const float dz = z > 0.f ? g[j] : 0.f;
        s_dz[j] += dz;
        s_dzx[j] = fmaf(dz, f[j], s_dzx[j]);
      }
    }
  }
  {
    float* mine = s_part + (size_t)lane_row * 2 * k.c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[v * 8 + j] = s_dz[j];
      mine[k.c + v * 8 + j] = s_dzx[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * k.c; i += kThreads) {
    float tot = 0.f;
    for (int r = 0; r < rows_per_blk; ++r) tot += s_part[(size_t)r * 2 * k.c + i];
    if (accum != nullptr) atomicAdd(accum + i, (double)tot);
    else partials[(size_t)blockIdx.x * 2 * k.c + i] = tot;
  }
}

__global__ void __launch_bounds__(32 * kFinGroups)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int num_partials, int channels, double count,
                       const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ coef, int accumulate) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  __shared__ double s_a[kFinGroups][32], s_b[kFinGroups][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double a = 0.0, b = 0.0;
  if (c < channels) {
    double a4[4] = {0.0, 0.0, 0.0, 0.0}, b4[4] = {0.0, 0.0, 0.0, 0.0};
    int p = grp;
    for (; p + 3 * kFinGroups < num_partials; p += 4 * kFinGroups) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        x[u] = partials[((size_t)(p + u * kFinGroups) * 2) * channels + c];
        y[u] = partials[((size_t)(p + u * kFinGroups) * 2 + 1) * channels + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a4[u] += (double)x[u]; b4[u] += (double)y[u]; }
    }
    for (; p < num_partials; p += kFinGroups) {
      a4[0] += (double)partials[((size_t)p * 2) * channels + c];
      b4[0] += (double)partials[((size_t)p * 2 + 1) * channels + c];
    }
    a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    b = (b4[0] + b4[1]) + (b4[2] + b4[3]);
  }
  s_a[grp][lane] = a;
  s_b[grp][lane] = b;
  __syncthreads();
  if (grp == 0 && c < channels) {
    for (int g = 1; g < kFinGroups; ++g) { a += s_a[g][lane]; b += s_b[g][lane]; }
    b = (double)invstd[c] * (b - (double)mean[c] * a);        // sum(dz * y) -> sum(dz * xhat)
    if (accumulate) { dbeta[c] += (float)a; dgamma[c] += (float)b; }
    else { dbeta[c] = (float)a; dgamma[c] = (float)b; }
    coef[c] = gamma[c] * invstd[c];
    coef[channels + c] = (float)(a / count);
    coef[2 * channels + c] = (float)(b / count);
  }
}

// HAS_DB: also sum dy per channel (gradient of a bias in front of the BatchNorm).  That gradient is identically zero
// (sum over pixels of dy = c1 * (sum dz - N*mean(dz) - c3 * sum xhat) = 0): the fused Toy_Net leaves it at the zero
// the gradient arena is cleared to and runs the variant without the accumulation (8 registers, 32 adds per window,
// one global atomic per channel per CTA); the reference's numerical value is rounding noise of order 1e-7.
template <bool POOL, bool HAS_DB>
__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_apply_kernel(BwdCtx k, const float* __restrict__ coef, act_t* __restrict__ dy,
                    float* __restrict__ dbias, const BnFusedBwd fb) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  extern __shared__ float s_db[];                          // [c]
  __shared__ float s_c1[512], s_k0[512], s_k1[512];
  const int cv = k.c >> 3;
  const int v = threadIdx.x % cv, lane_row = threadIdx.x / cv, rows_per_blk = kThreads / cv;
  const int oh = POOL ? k.h >> 1 : k.h, ow = POOL ? k.w >> 1 : k.w;
  const long long total_pix = (long long)k.n * oh * ow;
  // dy = c1 * (dz - c2 - xhat * c3), xhat = (y - mean) * invstd   ==>   dy = c1 * dz + k0 - k1 * y
  float sc[8], sh[8], c1[8], k0[8], k1[8], db[8];
  load8f(k.scale + v * 8, sc); load8f(k.shift + v * 8, sh);
  if (fb.accum != nullptr) {
    // fused statistics: sum dz and sum dz*y arrive as fp64 accumulators; the first C/8 threads derive the three
    // per-channel constants (shared memory) and CTA 0 writes dgamma = sum dz*xhat and dbeta = sum dz
    if ((int)threadIdx.x < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = threadIdx.x * 8 + j;
        const double a = fb.accum[ch];
        const double is = (double)k.invstd[ch], mu = (double)k.mean[ch];
        const double b = is * (fb.accum[k.c + ch] - mu * a);           // sum(dz * y) -> sum(dz * xhat)
        const float cc1 = fb.gamma[ch] * k.invstd[ch];
        const float cc2 = (float)(a * fb.inv_count), cc3 = (float)(b * fb.inv_count);
        const float kk1 = cc1 * k.invstd[ch] * cc3;
        s_c1[ch] = cc1;
        s_k1[ch] = kk1;
        s_k0[ch] = fmaf(kk1, k.mean[ch], -cc1 * cc2);
        if (blockIdx.x == 0) { fb.dbeta[ch] = (float)a; fb.dgamma[ch] = (float)b; }
      }
    }
    __syncthreads();
    load8s(s_c1 + v * 8, c1); load8s(s_k0 + v * 8, k0); load8s(s_k1 + v * 8, k1);
#pragma unroll
    for (int j = 0; j < 8; ++j) db[j] = 0.f;
  } else {
  load8f(coef + v * 8, c1);
  {
    float mu[8], is[8], c2[8], c3[8];
    load8f(k.mean + v * 8, mu);  load8f(k.invstd + v * 8, is);
    load8f(coef + k.c + v * 8, c2); load8f(coef + 2 * k.c + v * 8, c3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k1[j] = c1[j] * is[j] * c3[j];
      k0[j] = fmaf(k1[j], mu[j], -c1[j] * c2[j]);
      db[j] = 0.f;
    }
  }
  }
  const long long stride = (long long)gridDim.x * rows_per_blk;
  if constexpr (POOL) {
    constexpr int U = 2;                                   // windows (4 + 1 16-byte loads each) in flight
    const PoolConsts pc = pool_consts(sc, sh);
    const unsigned total32 = (unsigned)total_pix, stride32 = (unsigned)stride;
    const unsigned first = blockIdx.x * rows_per_blk + lane_row;
    const unsigned plane = (unsigned)ow * oh;
    int pw[U], ph[U], img[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = first + u * stride32;
      pw[u] = (int)(p % ow); ph[u] = (int)((p / ow) % oh); img[u] = (int)(p / plane);
    }
    const unsigned step = U * stride32;
    const int d_pw = (int)(step % ow), d_ph = (int)((step / ow) % oh), d_img = (int)(step / plane);
    for (unsigned pix = first; pix < total32; pix += step) {
      uint4 yq[U][4], gq[U];
      size_t base_off[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pix + u * stride32 >= total32) continue;
        base_off[u] = (((size_t)img[u] * k.h + 2 * ph[u]) * k.w + 2 * pw[u]) * k.c + v * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) yq[u][q] = ldg16(k.y + base_off[u] + ((q >> 1) * k.w + (q & 1)) * k.c);
        gq[u] = load_g_packed(k, img[u], ph[u], pw[u], oh, ow, v);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pix + u * stride32 >= total32) continue;
        // routed gradient of the four window positions, still packed: the first position equal to the window
        // maximum receives g (if its z > 0), the others 0 -- the same tie rule as max_pool2d's backward
        uint32_t dz[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t sg = pc.sgn[i];
          const uint32_t a0 = comp(yq[u][0], i) ^ sg, a1 = comp(yq[u][1], i) ^ sg;
          const uint32_t a2 = comp(yq[u][2], i) ^ sg, a3 = comp(yq[u][3], i) ^ sg;
          const uint32_t best = bmax2(bmax2(a0, a1), bmax2(a2, a3));
          const uint32_t gg = comp(gq[u], i) & __hgt2_mask(as_bf162(best), as_bf162(pc.thr[i]));
          const uint32_t e0 = __heq2_mask(as_bf162(a0), as_bf162(best));
          const uint32_t e1 = __heq2_mask(as_bf162(a1), as_bf162(best));
          const uint32_t e2 = __heq2_mask(as_bf162(a2), as_bf162(best));
          dz[0][i] = gg & e0;
          dz[1][i] = gg & e1 & ~e0;
          dz[2][i] = gg & e2 & ~(e0 | e1);
          dz[3][i] = gg & ~(e0 | e1 | e2);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t yw = comp(yq[u][q], i);
            const float lo = fmaf(c1[2 * i], bf_lo(dz[q][i]), fmaf(-k1[2 * i], bf_lo(yw), k0[2 * i]));
            const float hi = fmaf(c1[2 * i + 1], bf_hi(dz[q][i]), fmaf(-k1[2 * i + 1], bf_hi(yw), k0[2 * i + 1]));
            if (HAS_DB) { db[2 * i] += lo; db[2 * i + 1] += hi; }
            o[i] = pack_act2(lo, hi);
          }
          *reinterpret_cast<uint4*>(dy + base_off[u] + ((q >> 1) * k.w + (q & 1)) * k.c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pw[u] += d_pw; if (pw[u] >= ow) { pw[u] -= ow; ++ph[u]; }
        ph[u] += d_ph; if (ph[u] >= oh) { ph[u] -= oh; ++img[u]; }
        img[u] += d_img;
      }
    }
  } else if (!k.flatten) {
    constexpr int U = 4;
    for (long long pix0 = (long long)blockIdx.x * rows_per_blk + lane_row; pix0 < total_pix; pix0 += stride * U) {
      uint4 yv[U], gv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pix = pix0 + u * stride;
        if (pix < total_pix) {
          yv[u] = ldg16(k.y + pix * k.c + v * 8);
          gv[u] = ldg16(k.g + pix * k.c + v * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pix = pix0 + u * stride;
        if (pix >= total_pix) continue;
        float f[8], g[8], o[8];
        unpack8(yv[u], f);
        unpack8(gv[u], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(f[j], sc[j], sh[j]);
          const float dz = z > 0.f ? g[j] : 0.f;
          o[j] = fmaf(c1[j], dz, fmaf(-k1[j], f[j], k0[j]));
          if (HAS_DB) db[j] += o[j];
        }
        *reinterpret_cast<uint4*>(dy + pix * k.c + v * 8) = pack8(o);
      }
    }
  } else {
    const int plane = oh * ow;
    for (long long pix = (long long)blockIdx.x * rows_per_blk + lane_row; pix < total_pix; pix += stride) {
      const int img = (int)(pix / plane), rem = (int)(pix % plane);
      float f[8], g[8], o[8];
      unpack8(ldg16(k.y + pix * k.c + v * 8), f);
      load_g(k, img, rem / ow, rem % ow, oh, ow, v, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(f[j], sc[j], sh[j]);
        const float dz = z > 0.f ? g[j] : 0.f;
        o[j] = fmaf(c1[j], dz, fmaf(-k1[j], f[j], k0[j]));
        if (HAS_DB) db[j] += o[j];
      }
      *reinterpret_cast<uint4*>(dy + pix * k.c + v * 8) = pack8(o);
    }
  }
  if constexpr (HAS_DB) {
    if (dbias == nullptr) return;
    for (int i = threadIdx.x; i < k.c; i += kThreads) s_db[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_db[v * 8 + j], db[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < k.c; i += kThreads) atomicAdd(dbias + i, s_db[i]);
  }
}

// ================================================================================================
// [N][C][P] -> [N][P][C] (bf16): the first Linear's data gradient arrives in torch's flatten order (c, h, w); the
// pooled BatchNorm backward wants channels innermost.  One image per CTA through shared memory, 16-byte global
// accesses on both sides.  ncu on the gather it replaces (8 strided 2-byte loads per window inside the BatchNorm
// passes): 49.6 + 63.4 us at 1.7 TB/s for conv.10, long-scoreboard bound (profiles/ncu_bnbwd_v9.txt).
// ================================================================================================
__global__ void __launch_bounds__(kThreads)
chw_to_hwc_kernel(const act_t* __restrict__ in, act_t* __restrict__ out, int n, int c, int p) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  extern __shared__ act_t s_tile[];                // [c][p + 2]
  const int pitch = p + 2;
  const int vec_per_img = c * p / 8;
  for (int img = blockIdx.x; img < n; img += gridDim.x) {
    const act_t* src = in + (size_t)img * c * p;
    for (int i = threadIdx.x; i < vec_per_img; i += kThreads) {
      const uint4 v = ldg16(src + (size_t)i * 8);
      const act_t* e = reinterpret_cast<const act_t*>(&v);
      const int flat = i * 8, ch = flat / p, pos = flat - ch * p;   // p % 8 == 0: one vector stays in one channel
#pragma unroll
      for (int j = 0; j < 8; ++j) s_tile[ch * pitch + pos + j] = e[j];
    }
    __syncthreads();
    act_t* dst = out + (size_t)img * c * p;
    const int cv = c / 8;
    for (int i = threadIdx.x; i < vec_per_img; i += kThreads) {
      const int pos = i / cv, v = i - pos * cv;
      act_t e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] = s_tile[(v * 8 + j) * pitch + pos];
      *reinterpret_cast<uint4*>(dst + (size_t)pos * c + v * 8) = *reinterpret_cast<const uint4*>(e);
    }
    __syncthreads();
  }
}

// ================================================================================================
// Head: Linear(512,10) + CE(mean)/B + argmax (+ backward)
// ================================================================================================
constexpr int kFeat = 512, kCls = 10, kHeadRows = 8;   // one warp per row, 8 rows per block iteration
constexpr int kHeadPartialCols = kCls * kFeat + kCls + 2;

__global__ void __launch_bounds__(kThreads)
head_kernel(HeadArgs a) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  __shared__ float s_w[kCls][kFeat];
  __shared__ float s_dw[kCls][kFeat];
  __shared__ float s_db[kCls];
  __shared__ float s_metric[2];
  const bool train = a.da != nullptr;
  for (int i = threadIdx.x; i < kCls * kFeat; i += kThreads) {
    (&s_w[0][0])[i] = a.w[i];
    if (train) (&s_dw[0][0])[i] = 0.f;
  }
  if (threadIdx.x < kCls) s_db[threadIdx.x] = 0.f;
  if (threadIdx.x < 2) s_metric[threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_b = 1.0f / (float)a.batch;
  const float gscale = train ? (a.loss_scale ? *a.loss_scale : 1.0f) * inv_b / a.extra_div : 0.f;
  float my_loss = 0.f, my_correct = 0.f;
  // uniform trip count (the ordered weight-gradient fold below synchronises the block every iteration)
  for (int base = blockIdx.x * kHeadRows; base < a.batch; base += gridDim.x * kHeadRows) {
    const int row = base + warp;
    const bool active = row < a.batch;
    float dl[kCls], x[16];
    if (active) {
    // lane owns features {64*j + 2*lane + e : j < 8, e < 2}: 4-byte coalesced loads, conflict-light smem
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const act2_t v = *reinterpret_cast<const act2_t*>(static_cast<const act_t*>(a.a) + (size_t)row * kFeat + 64 * j + 2 * lane);
      const float2 f = act2_to_float2(v);
      x[2 * j] = f.x;
      x[2 * j + 1] = f.y;
    }
    float logit[kCls];
#pragma unroll
    for (int k = 0; k < kCls; ++k) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) p = fmaf(x[j], s_w[k][64 * (j >> 1) + 2 * lane + (j & 1)], p);
      logit[k] = warp_sum(p) + a.b[k];
    }
    float mx = logit[0];
    int arg = 0;
#pragma unroll
    for (int k = 1; k < kCls; ++k) if (logit[k] > mx) { mx = logit[k]; arg = k; }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < kCls; ++k) den += __expf(logit[k] - mx);
    const int tgt = a.target != nullptr ? (int)a.target[row] : -1;
    float tl = logit[0];
#pragma unroll
    for (int k = 1; k < kCls; ++k) tl = (k == tgt) ? logit[k] : tl;
    if (lane == 0) {
      if (tgt >= 0) {
        my_loss += (mx + __logf(den)) - tl;
        my_correct += (arg == tgt) ? 1.f : 0.f;
      }
      if (a.logits != nullptr) {
#pragma unroll
        for (int k = 0; k < kCls; ++k) a.logits[(size_t)row * kCls + k] = logit[k];
      }
    }
    if (train) {
      const float inv_den = 1.0f / den;
#pragma unroll
      for (int k = 0; k < kCls; ++k)
        dl[k] = a.dlogits != nullptr ? a.dlogits[(size_t)row * kCls + k]
                                     : (__expf(logit[k] - mx) * inv_den - (k == tgt ? 1.f : 0.f)) * gscale;
      float dx[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) dx[j] = 0.f;
#pragma unroll
      for (int k = 0; k < kCls; ++k) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dx[j] = fmaf(dl[k], s_w[k][64 * (j >> 1) + 2 * lane + (j & 1)], dx[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint32_t*>(static_cast<act_t*>(a.da) + (size_t)row * kFeat + 64 * j + 2 * lane) = pack_act2(dx[2 * j], dx[2 * j + 1]);
    }
    }  // active
    if (train) {
      // dW += dlogits^T x, one warp (= one row) at a time in warp order: lanes own disjoint features, so these are
      // plain read-modify-writes and the sum over rows has a fixed association (no shared-memory float atomics)
      for (int w = 0; w < kHeadRows; ++w) {
        if (warp == w && active) {
#pragma unroll
          for (int k = 0; k < kCls; ++k) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s_dw[k][64 * (j >> 1) + 2 * lane + (j & 1)] += dl[k] * x[j];
            if (lane == 0) s_db[k] += dl[k];
          }
        }
        __syncthreads();
      }
    }
  }
  __shared__ float s_m[kHeadRows][2];
  if (lane == 0) { s_m[warp][0] = my_loss; s_m[warp][1] = my_correct; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadRows; ++w) { l += s_m[w][0]; c += s_m[w][1]; }
    s_metric[0] = l * inv_b / a.extra_div * (float)a.batch;             // batch_loss * B
    s_metric[1] = c;
  }
  __syncthreads();
  if (a.partials != nullptr) {
    // deterministic mode: one row per block [5120 dW | 10 db | 2 metrics]; reduce_rows_kernel folds them in block order
    float* row = a.partials + (size_t)blockIdx.x * kHeadPartialCols;
    if (train) {
      for (int i = threadIdx.x; i < kCls * kFeat; i += kThreads) row[i] = (&s_dw[0][0])[i];
      if (threadIdx.x < kCls) row[kCls * kFeat + threadIdx.x] = s_db[threadIdx.x];
    }
    if (threadIdx.x < 2) row[kCls * kFeat + kCls + threadIdx.x] = s_metric[threadIdx.x];
    return;
  }
  if (threadIdx.x == 0 && a.metrics != nullptr) {
    atomicAdd(a.metrics, s_metric[0]);
    atomicAdd(a.metrics + 1, s_metric[1]);
  }
  if (train) {
    for (int i = threadIdx.x; i < kCls * kFeat; i += kThreads) atomicAdd(a.dw + i, (&s_dw[0][0])[i]);
    if (threadIdx.x < kCls) atomicAdd(a.db + threadIdx.x, s_db[threadIdx.x]);
  }
}

// w [c_out][taps][c_in] (the K-major forward weight) -> wt [c_in][taps][c_out]: the K-major operand of the data
// gradient (reduction over c_out), which lets the 64-column data gradient run as a CTA pair.  144 KiB for conv.3.
__global__ void __launch_bounds__(kThreads)
transpose_taps_kernel(const act_t* __restrict__ w, act_t* __restrict__ wt, int c_out, int taps, int c_in) {
  pdl_wait();
  pdl_trigger();
  const int total = c_out * taps * c_in;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
    const int co = i % c_out, t = (i / c_out) % taps, ci = i / (c_out * taps);       // i indexes wt
    wt[i] = w[((size_t)co * taps + t) * c_in + ci];
  }
}

}  // namespace

void launch_transpose_taps(const act_t* w, act_t* wt, int c_out, int taps, int c_in, cudaStream_t s) {
  const int total = c_out * taps * c_in;
  int grid = (total + kThreads - 1) / kThreads;
  if (grid > 4 * sm_count()) grid = 4 * sm_count();
  launch_pdl(transpose_taps_kernel, dim3(grid), dim3(kThreads), 0, s, w, wt, c_out, taps, c_in);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

// ---- launchers -------------------------------------------------------------------------------------
#ifndef DDPT_ACT_F16
int conv0_partial_rows() { return sm_count() * 2; }

#endif
int launch_conv0_fwd(const void* x, int x_is_u8, const float* weight, const float* bias, act_t* y,
                     float* stat_partials, double* stat_accum, int batch, cudaStream_t s) {
  const int grid = batch < conv0_partial_rows() ? batch : conv0_partial_rows();
  launch_pdl(conv0_fwd_kernel, dim3(grid), dim3(kThreads), 0, s, x, x_is_u8, weight, bias, y, stat_partials, stat_accum, batch);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return grid;
}

#ifndef DDPT_ACT_F16
int conv0_wgrad_partial_rows() { return sm_count() * 2; }

#endif
void launch_conv0_wgrad(const void* x, int x_is_u8, const act_t* dy, float* dw, float* db, float* partials,
                        int batch, const Conv0BnFuse& bn, cudaStream_t s, int channels_per_thread) {
  // the fused variant holds 40 more per-thread constants (177 registers): one resident CTA per SM, so one CTA per SM
  const int per_sm = bn.y != nullptr ? 1 : 2;
  int grid = batch < sm_count() * per_sm ? batch : sm_count() * per_sm;
  static const int env_ch = std::getenv("DDPT_CONV0_WGRAD_CH") != nullptr ? std::atoi(std::getenv("DDPT_CONV0_WGRAD_CH")) : 8;
  if ((channels_per_thread != 0 ? channels_per_thread : env_ch) == 4) {        // 4 channels per thread, one 16-warp CTA per SM
    grid = batch < sm_count() ? batch : sm_count();
    launch_pdl(conv0_wgrad4_kernel, dim3(grid), dim3(kThreads4), 0, s, x, x_is_u8, dy, dw, db, partials, batch, bn);
  } else {
    launch_pdl(conv0_wgrad_kernel, dim3(grid), dim3(kThreads), 0, s, x, x_is_u8, dy, dw, db, partials, batch, bn);
  }
  if (partials != nullptr) {          // deterministic mode: fold the block rows in order (dw: 576 columns, db: 64)
    launch_pdl(reduce_rows_kernel, dim3((kC0 * 9 + kThreads - 1) / kThreads), dim3(kThreads), 0, s,
               (const float*)partials, grid, kC0 * 10, kC0 * 9, dw, 1);
    launch_pdl(reduce_rows_kernel, dim3(1), dim3(kThreads), 0, s, (const float*)(partials + kC0 * 9), grid, kC0 * 10, kC0, db, 1);
  }
  DDPT_CUDA_CHECK(cudaGetLastError());
}

#ifndef DDPT_ACT_F16
void launch_bn_finalize(const float* partials, int num_partials, int channels, double count, const float* gamma,
                        const float* beta, float eps, float momentum, BnCoeffs out, float* running_mean,
                        float* running_var, long long* nbt, const float* conv_bias, cudaStream_t s) {
  launch_pdl(bn_finalize_kernel, dim3((channels + 31) / 32), dim3(32 * kFinGroups), 0, s, partials, num_partials, channels, count, gamma, beta,
                                                                eps, momentum, out, running_mean, running_var, nbt, conv_bias);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float eps, int channels, float* scale, float* shift,
                           const float* conv_bias, cudaStream_t s) {
  launch_pdl(bn_eval_coeffs_kernel, dim3((channels + 127) / 128), dim3(128), 0, s, gamma, beta, running_mean, running_var, eps, channels,
                                                               scale, shift, conv_bias);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

#endif
static void check_channels(int c) {
  if (c % 8 != 0 || kThreads % (c / 8) != 0 || c > 2048)
    throw std::runtime_error("toynet elementwise kernels need C % 8 == 0 and (C/8) dividing 256");
}

void launch_bn_relu_pool(const act_t* y, const float* scale, const float* shift, act_t* out, int n,
                         int h, int w, int c, int pool, int flatten_chw, uint8_t* q_out, uint8_t* q_scale,
                         const BnFusedFwd& fused, cudaStream_t s) {
  check_channels(c);
  if (fused.accum != nullptr && c > 512) throw std::runtime_error("fused BatchNorm statistics support C <= 512");
  if (pool && ((h | w) & 1)) throw std::runtime_error("2x2 max-pool needs even H and W");
  if (flatten_chw) {
    if (!pool) throw std::runtime_error("flatten_chw is only implemented together with the pool");
    const int ppix = (h / 2) * (w / 2);
    const size_t smem = (size_t)c * (ppix + 2) * sizeof(act_t);
    launch_pdl(bn_relu_pool_flatten_kernel, dim3(n), dim3(kThreads), smem, s, y, scale, shift, out, h, w, c, q_out, q_scale, fused);
  } else {
    const long long pixels = (long long)n * (pool ? (h / 2) * (w / 2) : h * w);
    const int rows_per_blk = kThreads / (c / 8);
    const int grid = persistent_grid(pixels, rows_per_blk * (pool ? 2 : 4), 8);
    if (q_out != nullptr && (c % 32 != 0)) throw std::runtime_error("MXFP8 side output needs C % 32 == 0");
    if (pool) launch_pdl(bn_relu_pool_kernel<true>, dim3(grid), dim3(kThreads), 0, s, y, scale, shift, out, n, h, w, c, q_out, q_scale, fused);
    else launch_pdl(bn_relu_pool_kernel<false>, dim3(grid), dim3(kThreads), 0, s, y, scale, shift, out, n, h, w, c, q_out, q_scale, fused);
  }
  DDPT_CUDA_CHECK(cudaGetLastError());
}

#ifndef DDPT_ACT_F16
int bn_bwd_partial_rows() { return sm_count() * 4; }

#endif
int launch_bn_bwd_reduce(const act_t* y, const act_t* g, const float* scale, const float* shift,
                         const float* mean, const float* invstd, float* partials, double* accum, int n, int h, int w,
                         int c, int pool, int g_flatten_chw, cudaStream_t s) {
  check_channels(c);
  BwdCtx k{y, g, scale, shift, mean, invstd, n, h, w, c, g_flatten_chw};
  const long long pix = (long long)n * (pool ? (h / 2) * (w / 2) : h * w);
  const int rows_per_blk = kThreads / (c / 8);
  long long want = (pix + rows_per_blk - 1) / rows_per_blk;
  // one resident wave (2 CTAs per SM at 128 registers): half the partial rows for bn_bwd_finalize to fold
  const int cap = sm_count() * 2;
  const int grid = (int)(want < cap ? want : cap);
  const size_t smem = (size_t)rows_per_blk * 2 * c * sizeof(float);       // 16 KiB for any C
  if (pool) launch_pdl(bn_bwd_reduce_kernel<true>, dim3(grid), dim3(kThreads), smem, s, k, partials, accum);
  else launch_pdl(bn_bwd_reduce_kernel<false>, dim3(grid), dim3(kThreads), smem, s, k, partials, accum);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return grid;
}

#ifndef DDPT_ACT_F16
void launch_bn_bwd_finalize(const float* partials, int num_partials, int channels, double count, const float* gamma,
                            const float* mean, const float* invstd, float* dgamma, float* dbeta, float* coef,
                            int accumulate, cudaStream_t s) {
  launch_pdl(bn_bwd_finalize_kernel, dim3((channels + 31) / 32), dim3(32 * kFinGroups), 0, s, partials, num_partials, channels, count, gamma,
                                                                    mean, invstd, dgamma, dbeta, coef, accumulate);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

#endif
void launch_bn_bwd_apply(const act_t* y, const act_t* g, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* coef, act_t* dy, float* dbias,
                         const BnFusedBwd& fused, int n, int h, int w, int c, int pool, int g_flatten_chw,
                         cudaStream_t s) {
  check_channels(c);
  if (fused.accum != nullptr && c > 512) throw std::runtime_error("fused BatchNorm statistics support C <= 512");
  BwdCtx k{y, g, scale, shift, mean, invstd, n, h, w, c, g_flatten_chw};
  const long long pix = (long long)n * (pool ? (h / 2) * (w / 2) : h * w);
  const int rows_per_blk = kThreads / (c / 8);
  long long want = (pix + rows_per_blk - 1) / rows_per_blk;
  // CTAs per SM in the grid (2 are resident at a time).  Same-box A/B, 300 steps each: 8 per SM 1.568 / 1.560 ms,
  // 2 per SM (one wave) 1.573 / 1.572 ms -- short CTAs leave gaps the side-stream weight gradients fill.
  // DDPT_BN_APPLY_CTAS overrides.
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = std::getenv("DDPT_BN_APPLY_CTAS");
    per_sm = e != nullptr && std::atoi(e) > 0 ? std::atoi(e) : 8;
  }
  const int cap = sm_count() * per_sm;
  const int grid = (int)(want < cap ? want : cap);
  const size_t smem = c * sizeof(float);
  if (dbias != nullptr) {
    if (pool) launch_pdl(bn_bwd_apply_kernel<true, true>, dim3(grid), dim3(kThreads), smem, s, k, coef, dy, dbias, fused);
    else launch_pdl(bn_bwd_apply_kernel<false, true>, dim3(grid), dim3(kThreads), smem, s, k, coef, dy, dbias, fused);
  } else {
    if (pool) launch_pdl(bn_bwd_apply_kernel<true, false>, dim3(grid), dim3(kThreads), smem, s, k, coef, dy, dbias, fused);
    else launch_pdl(bn_bwd_apply_kernel<false, false>, dim3(grid), dim3(kThreads), smem, s, k, coef, dy, dbias, fused);
  }
  DDPT_CUDA_CHECK(cudaGetLastError());
}

#ifndef DDPT_ACT_F16
int head_partial_cols() { return kHeadPartialCols; }

#endif
#ifdef DDPT_ACT_F16
void launch_head_f16(const HeadArgs& args, cudaStream_t s) {
#else
void launch_head(const HeadArgs& args, cudaStream_t s) {
#endif
  int grid = (args.batch + kHeadRows - 1) / kHeadRows;
  const int cap = sm_count();
  if (grid > cap) grid = cap;
  launch_pdl(head_kernel, dim3(grid), dim3(kThreads), 0, s, args);
  if (args.partials != nullptr) {      // deterministic mode: ordered fold of the per-block rows
    const float* p = args.partials;
    if (args.da != nullptr) {
      launch_pdl(reduce_rows_kernel, dim3((kCls * kFeat + kThreads - 1) / kThreads), dim3(kThreads), 0, s, p, grid,
                 kHeadPartialCols, kCls * kFeat, args.dw, 1);
      launch_pdl(reduce_rows_kernel, dim3(1), dim3(kThreads), 0, s, p + kCls * kFeat, grid, kHeadPartialCols, kCls, args.db, 1);
    }
    if (args.metrics != nullptr)
      launch_pdl(reduce_rows_kernel, dim3(1), dim3(kThreads), 0, s, p + kCls * kFeat + kCls, grid, kHeadPartialCols, 2,
                 args.metrics, 1);
  }
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_chw_to_hwc(const act_t* in, act_t* out, int n, int c, int p, cudaStream_t s) {
  if (p % 8 != 0 || c % 8 != 0) throw std::runtime_error("chw_to_hwc: C and H*W must be multiples of 8");
  const size_t smem = (size_t)c * (p + 2) * sizeof(act_t);
  if (smem > 48 * 1024) throw std::runtime_error("chw_to_hwc: one image does not fit 48 KiB of shared memory");
  const int grid = n < 4 * sm_count() ? n : 4 * sm_count();
  launch_pdl(chw_to_hwc_kernel, dim3(grid), dim3(kThreads), smem, s, in, out, n, c, p);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
