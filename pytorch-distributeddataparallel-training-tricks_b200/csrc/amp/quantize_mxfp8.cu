// MXFP8 quantisation (the AMP "O1-fp8" path): e4m3 payload + one UE8M0 power-of-two scale per 32 elements.
// One warp handles 32 blocks at a time: lane L owns block L of a 1024-element span (32 consecutive elements,
// four 16-byte loads), computes the block amax, derives the shared exponent and writes 32 fp8 bytes + 1 scale.
#include "../include/ddpt_fp8.h"
#include "../include/ddpt_common.cuh"
#include <cuda_fp8.h>

namespace ddpt {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint8_t to_e4m3(float v) {
  return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
}

template <bool SRC_FP32>
__global__ void __launch_bounds__(kThreads)
quantize_mxfp8_kernel(const void* __restrict__ src, uint8_t* __restrict__ data, uint8_t* __restrict__ scale,
                      long long num_blocks) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  for (long long blk = (long long)blockIdx.x * kThreads + threadIdx.x; blk < num_blocks;
       blk += (long long)gridDim.x * kThreads) {
    float v[32];
    if (SRC_FP32) {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(src) + blk * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(p + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    } else {
      const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(src) + blk * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 t = __ldg(p + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          v[8 * i + 2 * j] = f.x;
          v[8 * i + 2 * j + 1] = f.y;
        }
      }
    }
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(v[i]));
    // smallest power of two s with amax / s <= 448 (e4m3 max); s = 2^(e - 127)
    int e = 127;
    if (amax > 0.f && is_finite_f32(amax)) {
      int ex;
      const float m = frexpf(amax * (1.0f / 448.0f), &ex);   // amax/448 = m * 2^ex, m in [0.5, 1)
      e = 127 + ex - (m == 0.5f ? 1 : 0);
      e = max(1, min(254, e));
    }
    const float inv = exp2f(static_cast<float>(127 - e));
    uint32_t packed[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      packed[i] = (uint32_t)to_e4m3(v[4 * i] * inv) | ((uint32_t)to_e4m3(v[4 * i + 1] * inv) << 8) |
                  ((uint32_t)to_e4m3(v[4 * i + 2] * inv) << 16) | ((uint32_t)to_e4m3(v[4 * i + 3] * inv) << 24);
    uint4* o = reinterpret_cast<uint4*>(data + blk * 32);
    o[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    o[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    scale[blk] = static_cast<uint8_t>(e);
  }
}

}  // namespace

void launch_quantize_mxfp8(const void* src, int src_is_fp32, uint8_t* data_out, uint8_t* scale_out, long long rows,
                           int cols, cudaStream_t stream) {
  const long long num_blocks = rows * (cols / 32);
  if (num_blocks <= 0) return;
  long long want = (num_blocks + kThreads - 1) / kThreads;
  const int grid = (int)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
  if (src_is_fp32) launch_pdl(quantize_mxfp8_kernel<true>, dim3(grid), dim3(kThreads), 0, stream, src, data_out, scale_out, num_blocks);
  else launch_pdl(quantize_mxfp8_kernel<false>, dim3(grid), dim3(kThreads), 0, stream, src, data_out, scale_out, num_blocks);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
