// Shared device helpers for the ddpt_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace ddpt {

#define DDPT_CUDA_CHECK(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::ddpt::report_cuda_error(_e, #expr, __FILE__, __LINE__);                        \
    }                                                                                  \
  } while (0)

void report_cuda_error(cudaError_t e, const char* expr, const char* file, int line);
int sm_count();  // cached multiprocessor count of the current device

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// Every kernel of the training step is launched with the programmatic-stream-serialization attribute and starts with
// pdl_wait(): its CTAs may be scheduled (and run their parameter-only prologue: barrier init, TMEM allocation,
// tensor-map prefetch) while the previous kernel of the stream drains, but touch global memory only after that
// kernel has completed and flushed.  pdl_trigger() lets the NEXT kernel do the same.  Inside a CUDA graph the
// edges become programmatic dependencies.  DDPT_PDL=0 launches without the attribute (then both are no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DDPT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
#endif

// Loss-scaler state block shared by amp / all-reduce / optimizer kernels:
//   [0] scale  [1] 1/scale  [2] found_inf (0/1)  [3] clean steps  [4] skipped steps
enum ScalerSlot { kScale = 0, kInvScale = 1, kFoundInf = 2, kCleanSteps = 3, kSkipped = 4 };

__device__ __forceinline__ bool is_finite_f32(float v) {
  return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
}

__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace ddpt
