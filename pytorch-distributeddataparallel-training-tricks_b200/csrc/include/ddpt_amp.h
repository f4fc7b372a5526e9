// Mixed-precision support kernels (Apex amp_C equivalents) -- host API.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ddpt {

constexpr int kMaxTensorsPerLaunch = 48;

struct TensorList {
  int count;
  long long numel[kMaxTensorsPerLaunch];
  const void* in0[kMaxTensorsPerLaunch];
  const void* in1[kMaxTensorsPerLaunch];
  void* out[kMaxTensorsPerLaunch];
};

// out = in0 * (*scale_ptr) ; raises scaler_state[kFoundInf] on inf/nan.   (amp_C.multi_tensor_scale)
void launch_multi_tensor_scale(const TensorList& list, const float* scale_ptr, float* found_inf,
                               cudaStream_t stream);
// out = a*in0 + b*in1 ; raises found_inf on inf/nan in the selected argument(s). (amp_C.multi_tensor_axpby)
void launch_multi_tensor_axpby(const TensorList& list, float a, float b, int arg_to_check, float* found_inf,
                               cudaStream_t stream);
// Flat-arena variant: grads[0..n) *= scaler[kInvScale], flag -> scaler[kFoundInf].
void launch_unscale_arena(float* grads, long long n, float* scaler_state, cudaStream_t stream);
// Dynamic loss-scale bookkeeping entirely on device (no host read of the overflow flag).
void launch_update_scale(float* scaler_state, int growth_interval, float min_scale, float max_scale,
                         cudaStream_t stream);

}  // namespace ddpt
