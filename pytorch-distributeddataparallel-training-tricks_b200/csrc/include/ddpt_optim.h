// Fused Lookahead + Nesterov-SGD update over flat arenas -- host API.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ddpt {

// Device-resident hyper-parameters (float[8]) so CUDA-graph replays see scheduler changes.
enum HyperSlot { kLr = 0, kMomentum = 1, kWeightDecay = 2, kAlpha = 3, kDampening = 4, kHyperSlots = 8 };
// Device-resident integer state (int[8]).
enum OptStateSlot { kStepCount = 0, kLaCounter = 1, kSlowReady = 2, kLaK = 3, kTicket = 4, kSkippedSteps = 5,
                    kOptStateSlots = 8 };

struct FusedSgdArgs {
  float* params;                 // fp32 master weights (arena)
  const float* grads;            // fp32 gradients (arena, already unscaled)
  float* momentum;               // fp32 momentum buffers (arena)
  float* slow;                   // fp32 Lookahead slow weights (arena) or nullptr (no Lookahead)
  __nv_bfloat16* shadow;         // optional 16-bit copy of the weights for the tensor-core path
  int shadow_fp16;               // 0: the shadow is bf16, 1: fp16 (O1-fp16)
  float* grads_to_zero;          // optional: zero the gradient arena after it has been consumed
  const float* hyper;            // float[kHyperSlots]
  int* state;                    // int[kOptStateSlots]
  const float* found_inf;        // optional scaler flag: non-zero => the whole step is skipped
  long long numel;               // arena length (multiple of 4)
  int nesterov;
};

void launch_fused_lookahead_sgd(const FusedSgdArgs& args, cudaStream_t stream);
// fp32 -> bf16 re-cast of an arena (used once at attach time and after load_state_dict).
void launch_cast_bf16(const float* src, __nv_bfloat16* dst, long long numel, cudaStream_t stream, int fp16 = 0);

}  // namespace ddpt
