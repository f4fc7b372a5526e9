// conv.0 weight gradient (1 -> 64 channels, 3x3; reference model.py:17) on the tensor cores, with block 0's
// BatchNorm-backward "apply" folded into the operand producer.
//
//     dW[ch, tap] = sum over (image, pixel) of dy[pixel, ch] * x[pixel + tap]          db[ch] = sum dy[pixel, ch]
//
// The CUDA-core version keeps 80 fp32 accumulators per thread (8 channels x (9 taps + bias)): 80 FFMA per 16 bytes of
// gradient, ~210 registers, one CTA of 8 warps per SM -- issue bound at a third of the copy bandwidth
// (profiles/ncu_r2_conv0_wgrad.txt).  As a GEMM it is  D[64 ch, N] = dy^T[64 ch, K pixels] . P[K pixels, N]  with
// P = [9 taps of x | 1 | 9 taps of (x - round16(x)) | 0...]: one 128 x 32 x 16 tcgen05.mma per 16 pixels.  The threads
// only produce operands:
//   A  (MN-major, 64 channels x 128 pixels per stage): dy = c1 * dz + k0 - k1 * y0 computed from the saved BatchNorm
//      input y0 and the incoming gradient g (or dy itself when not fused), rounded to the activation type, written with
//      the 128-byte swizzle -- the NHWC row of a pixel IS an MN-major operand row;
//   B  (K-major, 32 rows x 128 pixels in two 64-pixel blocks): the image patch values; the second group of 9 rows carries the rounding
//      remainder of x so a float image loses nothing to the 16-bit operand (uint8 images are exact: raw 0..255 in
//      the operand, 1/255 applied to the result).
// The accumulator lives in TMEM for the whole kernel (M = 128 rows: rows 64..127 multiply whatever bytes follow the
// stage and are never read).  Warp roles: 8 producer warps, 1 issuing warp; warps 0 / 1 read the result at the end.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <stdexcept>
#include "../include/ddpt_toynet.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"

namespace ddpt {
namespace {

using namespace umma;

constexpr int kImg = 28, kOut = 26, kPix = kOut * kOut, kC = 64;
constexpr int kProducers = 256;
constexpr int kThreadsTc = kProducers + 32;
constexpr int kStagePix = 128;                             // pixels (GEMM K) per stage: 4 gradient vectors (+ 4 of y0) per
                                                           // thread in flight -- the kernel is a latency-bound stream otherwise
constexpr int kPixPerThread = kStagePix / 32;
constexpr int kStagesPerImage = (kPix + kStagePix - 1) / kStagePix;     // 6 (the last one zero-padded)
constexpr int kABytes = kStagePix * 128;                   // 128 pixel rows x 64 channels
constexpr int kN = 32;
constexpr int kBBlock = kN * 128;                          // 32 rows x 64 pixels (one 128-byte row holds 64 K elements)
constexpr int kBBytes = (kStagePix / 64) * kBBlock;
constexpr int kStages = 3;
constexpr int kRowFloats = kC * 10;                        // [576 dW | 64 db] per deterministic partial row

template <bool F16> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
template <bool F16> __device__ __forceinline__ float2 unpack2(uint32_t v) {
  if (F16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <bool F16> __device__ __forceinline__ float round16(float v) {
  return F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16(v));
}
template <bool F16> __device__ __forceinline__ unsigned short bits16(float v) {
  if (F16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16(v));
}
__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory"); }

template <bool F16>
__global__ void __maxnreg__(112)                 // 288 threads: two CTAs (18 warps) per SM
conv0_wgrad_tc_kernel(const void* __restrict__ x, const int x_is_u8, const void* __restrict__ dy,
                      float* __restrict__ dw, float* __restrict__ db, float* __restrict__ partials, const int batch,
                      const Conv0BnFuse bn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // kStages x 16 KB
  uint8_t* smem_b = smem + kStages * kABytes;              // kStages x 8 KB  (also the "rows 64..127" of the last A stage)
  float* s_img = reinterpret_cast<float*>(smem_b + kStages * kBBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_img + kImg * kImg);
  uint64_t* empty = full + kStages;
  uint64_t* done = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], kProducers); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int my_images = blockIdx.x < batch ? (batch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t kIdesc = make_idesc_bf16(128, kN, true, false) ^ (F16 ? kIdescBf16ToF16 : 0u);

  if (warp == 8) {
    // ===== MMA issuer: eight K = 16 steps per 128-pixel stage, one accumulator for the whole kernel =====
    if (elect_one()) {
      const int total = my_images * kStagesPerImage;
      int s = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total; ++it) {
        mbar_wait(&full[s], phase, 71);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * kABytes), kABytes, 1024);
        const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * kBBytes), 0, 1024);
#pragma unroll
        for (int k = 0; k < kStagePix / 16; ++k)
          mma_bf16_ss(tmem_base, desc_advance(adesc, k * (16 * 128)),
                      desc_advance(bdesc, (k / 4) * kBBlock + (k % 4) * 32), kIdesc, (it | k) != 0 ? 1u : 0u);
        mma_commit(&empty[s]);
        if (++s == kStages) { s = 0; phase ^= 1; }
      }
      mma_commit(done);
    }
  } else {
    // ===== operand producers =====
    const int cg = threadIdx.x & 7, prow = threadIdx.x >> 3;       // A: 8 channels of pixels prow + 32 u
    const int brow = threadIdx.x >> 3, bchunk = threadIdx.x & 7;   // B: row brow, pixels (64 blk + 8 bchunk) .. + 7
    const bool fused = bn.y != nullptr;
    float f_sc[8], f_sh[8], f_c1[8], f_k0[8], f_k1[8];
    if (fused) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        const float c1 = __ldg(bn.bcoef + c), c2 = __ldg(bn.bcoef + kC + c), c3 = __ldg(bn.bcoef + 2 * kC + c);
        f_sc[j] = __ldg(bn.scale + c); f_sh[j] = __ldg(bn.shift + c);
        f_c1[j] = c1;
        f_k1[j] = c1 * __ldg(bn.invstd + c) * c3;
        f_k0[j] = fmaf(f_k1[j], __ldg(bn.mean + c), -c1 * c2);
      }
    }
    const int tap = brow < 9 ? brow : brow - 10;                   // rows 0..8: x, 9: ones, 10..18: remainder of x
    const int tap_off = (brow < 9 || (brow >= 10 && brow < 19)) ? (tap / 3) * kImg + tap % 3 : 0;
    int s = 0;
    uint32_t phase = 0;
    for (int img = blockIdx.x; img < batch; img += gridDim.x) {
      producer_sync();                                             // everyone is done with the previous image's patch
      if (x_is_u8) {
        const uint8_t* src = static_cast<const uint8_t*>(x) + (size_t)img * kImg * kImg;
        for (int i = threadIdx.x; i < kImg * kImg / 4; i += kProducers) {
          const uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
          s_img[4 * i] = v.x; s_img[4 * i + 1] = v.y; s_img[4 * i + 2] = v.z; s_img[4 * i + 3] = v.w;
        }
      } else {
        const float* src = static_cast<const float*>(x) + (size_t)img * kImg * kImg;
        for (int i = threadIdx.x; i < kImg * kImg; i += kProducers) s_img[i] = src[i];
      }
      producer_sync();
      for (int blk = 0; blk < kStagesPerImage; ++blk) {
        // request this stage's gradient vectors before waiting for the slot
        uint4 raw[kPixPerThread], rawy[kPixPerThread];
#pragma unroll
        for (int u = 0; u < kPixPerThread; ++u) {
          const int p = blk * kStagePix + prow + u * 32;
          raw[u] = make_uint4(0, 0, 0, 0); rawy[u] = make_uint4(0, 0, 0, 0);
          if (p < kPix) {
            const size_t off = (((size_t)img * kPix + p) * kC + cg * 8) * 2;
            if (fused) {
              raw[u] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(bn.g) + off));
              rawy[u] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(bn.y) + off));
            } else {
              raw[u] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(dy) + off));
            }
          }
        }
        // B chunks: 8 consecutive pixels of row brow in each 64-pixel block
        uint32_t bw[kStagePix / 64][4];
#pragma unroll
        for (int hb = 0; hb < kStagePix / 64; ++hb)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          unsigned short h[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int p = blk * kStagePix + hb * 64 + bchunk * 8 + 2 * i + e;
            float v = 0.f;
            if (brow == 9) {
              v = 1.f;
            } else if (brow < 19 && p < kPix) {
              const int oh = p / kOut, ow = p - oh * kOut;
              const float xv = s_img[oh * kImg + ow + tap_off];
              v = brow < 9 ? xv : xv - round16<F16>(xv);
            }
            h[e] = bits16<F16>(v);
          }
          bw[hb][i] = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
        }
        mbar_wait(&empty[s], phase ^ 1, 72);
        uint8_t* a_tile = smem_a + s * kABytes;
#pragma unroll
        for (int u = 0; u < kPixPerThread; ++u) {
          const int r = prow + u * 32;                              // pixel row of the stage
          uint4 o = raw[u];
          if (fused) {
            const uint32_t gw[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            const uint32_t yw[4] = {rawy[u].x, rawy[u].y, rawy[u].z, rawy[u].w};
            uint32_t ow4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 g2 = unpack2<F16>(gw[i]), y2 = unpack2<F16>(yw[i]);
              const float dz0 = fmaf(y2.x, f_sc[2 * i], f_sh[2 * i]) > 0.f ? g2.x : 0.f;
              const float dz1 = fmaf(y2.y, f_sc[2 * i + 1], f_sh[2 * i + 1]) > 0.f ? g2.y : 0.f;
              ow4[i] = pack2<F16>(fmaf(f_c1[2 * i], dz0, fmaf(-f_k1[2 * i], y2.x, f_k0[2 * i])),
                                  fmaf(f_c1[2 * i + 1], dz1, fmaf(-f_k1[2 * i + 1], y2.y, f_k0[2 * i + 1])));
            }
            const bool live = blk * kStagePix + r < kPix;           // padding pixels of the last stage contribute nothing
            o = live ? make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]) : make_uint4(0, 0, 0, 0);
          }
          *reinterpret_cast<uint4*>(a_tile + r * 128 + ((cg ^ (r & 7)) << 4)) = o;
        }
#pragma unroll
        for (int hb = 0; hb < kStagePix / 64; ++hb)
          *reinterpret_cast<uint4*>(smem_b + s * kBBytes + hb * kBBlock + brow * 128 + ((bchunk ^ (brow & 7)) << 4)) =
              make_uint4(bw[hb][0], bw[hb][1], bw[hb][2], bw[hb][3]);
        fence_proxy_async();                                        // generic-proxy stores -> visible to the tensor core
        mbar_arrive(&full[s]);
        if (++s == kStages) { s = 0; phase ^= 1; }
      }
    }
    // ===== result: TMEM lanes 0..63 = channels; columns = [9 taps | 1 | 9 remainder taps] =====
    if (warp < 2 && my_images > 0) {
      mbar_wait(done, 0, 73);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), v);
      tmem_ld_wait();
      const int ch = warp * 32 + lane;
      const float scale = x_is_u8 ? 1.0f / 255.0f : 1.0f;
      float* row = partials != nullptr ? partials + (size_t)blockIdx.x * kRowFloats : nullptr;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float g = (__uint_as_float(v[t]) + __uint_as_float(v[10 + t])) * scale;
        if (row != nullptr) row[ch * 9 + t] = g;
        else atomicAdd(dw + ch * 9 + t, g);
      }
      if (row != nullptr) row[kC * 9 + ch] = __uint_as_float(v[9]);
      else atomicAdd(db + ch, __uint_as_float(v[9]));
    } else if (warp < 2 && partials != nullptr) {
      float* row = partials + (size_t)blockIdx.x * kRowFloats;     // a block without images still owns a (zero) row
      for (int i = lane + warp * 32; i < kRowFloats; i += 64) row[i] = 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

// out[i] += sum over rows r (in order) of partials[r * ld + i]
__global__ void __launch_bounds__(256)
fold_rows_kernel(const float* __restrict__ partials, int rows, int ld, int n, float* __restrict__ dw, float* __restrict__ db) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * 256 + threadIdx.x;
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
  if (i < kC * 9) dw[i] += tot;
  else db[i - kC * 9] += tot;
}

}  // namespace

void launch_conv0_wgrad_tc(const void* x, int x_is_u8, const void* dy, int fp16, float* dw, float* db, float* partials,
                           int batch, const Conv0BnFuse& bn, cudaStream_t s) {
  const size_t smem = 1024 + kStages * (kABytes + kBBytes) + kImg * kImg * 4 + (2 * kStages + 1) * 8 + 16;
  const int grid = std::min(batch, 2 * sm_count());         // conv0_wgrad_partial_rows() rows at most
  auto run = [&](auto kernel) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(kernel, dim3(grid), dim3(kThreadsTc), smem, s, x, x_is_u8, dy, dw, db, partials, batch, bn);
  };
  if (fp16) run(conv0_wgrad_tc_kernel<true>);
  else run(conv0_wgrad_tc_kernel<false>);
  DDPT_CUDA_CHECK(cudaGetLastError());
  if (partials != nullptr) {
    launch_pdl(fold_rows_kernel, dim3((kRowFloats + 255) / 256), dim3(256), 0, s, (const float*)partials, grid, kRowFloats,
               kRowFloats, dw, db);
    DDPT_CUDA_CHECK(cudaGetLastError());
  }
}

}  // namespace ddpt
