// Bandwidth-bound kernels of the fused Toy_Net pipeline (NHWC bf16) -- host API.
// BatchNorm statistics flow as "partials": float [P][2][C] (row 0: sums, row 1: sums of squares
// or sums of dz*xhat), produced by conv epilogues / reduce kernels and folded by *_finalize.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace ddpt {

// ---- conv.0 : 1 -> 64 channels, 3x3, CUDA cores (K = 9 is hostile to tensor cores) -------------
// x: uint8 [B,28,28] (scaled by 1/255 on the fly) or fp32 [B,28,28];  y: bf16 [B,26,26,64]
// Returns the number of statistic partial rows written ([rows][2][64]); capacity needed: conv0_partial_rows().
int conv0_partial_rows();
// stat_accum (optional, fp64 [2][64], zeroed by the caller): the block sums are added there (red.global.add.f64)
// instead of being written as partial rows.
int launch_conv0_fwd(const void* x, int x_is_u8, const float* weight /*[64][9]*/, const float* bias,
                     __nv_bfloat16* y, float* stat_partials /*[rows][2][64] or null*/, double* stat_accum,
                     int batch, cudaStream_t s);
// dW [64][9], db [64] accumulated with atomics into zero-initialised (or beta=1) fp32 buffers.
// partials (optional, [conv0_wgrad_partial_rows()][640] fp32): deterministic mode -- every block writes one row and an
// ordered reduction kernel folds them into dw / db instead of global float atomics.
// bn (optional): fold block 0's BatchNorm-backward apply into this kernel -- dy is computed from the saved BatchNorm input
// y0, the incoming gradient g and the finalize coefficients instead of being read (conv.0 has no data gradient, so
// nothing else needs dy0: the apply pass and its 88 MB write + read disappear).
struct Conv0BnFuse {
  const void* y;               // null: read dy (activation type of the overload called: bf16 or fp16)
  const void* g;
  const float* scale; const float* shift; const float* mean; const float* invstd;
  const float* bcoef;          // [3][64] from bn_bwd_finalize
};
int conv0_wgrad_partial_rows();
// channels_per_thread: 8 (8 warps per SM, 80 accumulators per thread), 4 (16 warps, 40 accumulators) or 0 = the default
// (DDPT_CONV0_WGRAD_CH, else the measured-best one).
void launch_conv0_wgrad(const void* x, int x_is_u8, const __nv_bfloat16* dy, float* dw, float* db, float* partials,
                        int batch, const Conv0BnFuse& bn, cudaStream_t s, int channels_per_thread = 0);
// Same contract on the tensor cores (csrc/nn/conv0_wgrad_tc.cu): dW^T = dy^T . [patches | 1] as 128 x 32 x 16 tcgen05.mma
// over operands the threads write to shared memory (dy derived from y0 / g on the fly when bn.y is set); opt-in.
void launch_conv0_wgrad_tc(const void* x, int x_is_u8, const void* dy, int fp16, float* dw, float* db, float* partials,
                           int batch, const Conv0BnFuse& bn, cudaStream_t s);

// ---- BatchNorm (training) ---------------------------------------------------------------------------
struct BnCoeffs {          // all float [C]
  float* scale;            // gamma * invstd
  float* shift;            // beta - mean * scale
  float* mean;
  float* invstd;
};
// Statistics handed over as fp64 accumulators [2][C] (sum, sum of squares) filled by the producing kernel's epilogue:
// the consuming pass derives scale/shift in its prologue (every CTA, 8 channels per thread, into shared memory) and
// CTA 0 publishes the coefficients for the backward pass and updates the running statistics -- the bn_finalize
// launch and its pass over hundreds of partial rows disappear from the critical path.
struct BnFusedFwd {
  const double* accum;     // null: use the precomputed scale/shift arguments
  double count;            // elements per channel; the kernels only multiply (1/count and count/(count-1) come from the host:
  double inv_count;        // fp64 divisions / square roots in every CTA's prologue cost more than the launch they replace)
  float unbias;
  const float* gamma; const float* beta;
  float eps, momentum;
  BnCoeffs out;
  float* running_mean; float* running_var; long long* num_batches_tracked;   // optional
};
// Same for the backward pass: accumulators [2][C] = sum dz, sum dz*y from bn_bwd_reduce; bn_bwd_apply derives its
// coefficients and CTA 0 writes dgamma / dbeta.
struct BnFusedBwd {
  const double* accum;     // null: use the precomputed coefficient argument
  double count, inv_count;
  const float* gamma;
  float* dgamma; float* dbeta;
};
// Fold partials into batch statistics, produce scale/shift, update running stats (momentum, unbiased var)
// and num_batches_tracked (int64).
void launch_bn_finalize(const float* partials, int num_partials, int channels, double count, const float* gamma,
                        const float* beta, float eps, float momentum, BnCoeffs out, float* running_mean,
                        float* running_var, long long* num_batches_tracked, const float* conv_bias, cudaStream_t s);
// Eval mode: scale/shift from the running statistics.
void launch_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float eps, int channels, float* scale, float* shift,
                           const float* conv_bias, cudaStream_t s);
// a = relu(y * scale + shift), optionally followed by 2x2 max-pool; flatten_chw: write the pooled map
// as [N][C*PH*PW] in (c, h, w) order -- torch's Flatten of an NCHW tensor -- for the first Linear.
// q_out / q_scale (optional): also emit the activation as MXFP8 (e4m3 [same layout] + UE8M0 per 32 along the last dim)
void launch_bn_relu_pool(const __nv_bfloat16* y, const float* scale, const float* shift, __nv_bfloat16* out, int n,
                         int h, int w, int c, int pool, int flatten_chw, uint8_t* q_out, uint8_t* q_scale,
                         const BnFusedFwd& fused, cudaStream_t s);

// ---- BatchNorm backward (two passes) ----------------------------------------------------------------------
// Pass 1: partials[blk][0][c] = sum dz, partials[blk][1][c] = sum dz * y (finalize converts to sum dz * xhat), dz = grad of the BN output
// after routing g through max-pool (if any) and the ReLU mask.  Returns the number of partial rows.
int bn_bwd_partial_rows();
int launch_bn_bwd_reduce(const __nv_bfloat16* y, const __nv_bfloat16* g, const float* scale, const float* shift,
                         const float* mean, const float* invstd, float* partials, double* accum /*[2][C] or null*/,
                         int n, int h, int w, int c, int pool, int g_flatten_chw, cudaStream_t s);
// Fold pass-1 partials: dgamma = sum dz*xhat, dbeta = sum dz (written/accumulated into the grad arena) and the
// three per-channel coefficients of pass 2.
void launch_bn_bwd_finalize(const float* partials, int num_partials, int channels, double count, const float* gamma,
                            const float* mean, const float* invstd, float* dgamma, float* dbeta,
                            float* coef /*[3][C]*/, int accumulate, cudaStream_t s);
// [N][C][P] -> [N][P][C] bf16 (flatten order of the first Linear's data gradient -> channels-last)
void launch_chw_to_hwc(const __nv_bfloat16* in, __nv_bfloat16* out, int n, int c, int p, cudaStream_t s);
// Pass 2: dy = c1 * (dz - c2 - xhat * c3) as bf16 NHWC; optionally sum dy per channel (conv bias gradient).
void launch_bn_bwd_apply(const __nv_bfloat16* y, const __nv_bfloat16* g, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* coef, __nv_bfloat16* dy,
                         float* dbias /*[C] atomics or null*/, const BnFusedBwd& fused, int n, int h, int w, int c,
                         int pool, int g_flatten_chw, cudaStream_t s);

// ---- classifier head: Linear(512,10) + CrossEntropy(mean)/B + argmax, forward and backward ------------------
struct HeadArgs {
  const void* a;               // [B,512] relu(bn(y4)); bf16 (launch_head) or fp16 (launch_head_f16)
  const float* w;              // [10,512] fp32 master weights
  const float* b;              // [10]
  const long long* target;     // [B]
  int batch;
  float* logits;               // [B,10] fp32 (also what the caller returns as model output)
  float* metrics;              // [2]: += sum_i CE_i / B   and  += #correct
  // training only (null in eval)
  const float* loss_scale;     // device scalar (amp) or null => 1
  void* da;                    // [B,512] grad wrt a (before the ReLU mask), same type as a
  float* dw;                   // [10,512] += dlogits^T a
  float* db;                   // [10]     += sum dlogits
  const float* dlogits;        // optional [B,10]: externally supplied d(loss)/d(logits) (autograd path)
  float extra_div;             // Q1: additional division of the loss by B (1/B) -> pass (float)B, or 1
  float* partials;             // optional [#SMs][head_partial_cols()]: deterministic mode (ordered cross-block fold of
                               // dw / db / metrics instead of global float atomics)
};
int head_partial_cols();
void launch_head(const HeadArgs& args, cudaStream_t s);
void launch_head_f16(const HeadArgs& args, cudaStream_t s);

// [c_out][taps][c_in] -> [c_in][taps][c_out] (bf16): K-major weights for the data gradient
void launch_transpose_taps(const __nv_bfloat16* w, __nv_bfloat16* wt, int c_out, int taps, int c_in, cudaStream_t s);


// ---- fp16-activation overloads (O1-fp16: Apex O1's precision in the reference); same semantics, same launch shapes.
// Defined by csrc/nn/toynet_elementwise_f16.cu, which rebuilds toynet_elementwise.cu with act_t = __half.
int launch_conv0_fwd(const void* x, int x_is_u8, const float* weight /*[64][9]*/, const float* bias,
                     __half* y, float* stat_partials /*[rows][2][64] or null*/, double* stat_accum,
                     int batch, cudaStream_t s);
void launch_conv0_wgrad(const void* x, int x_is_u8, const __half* dy, float* dw, float* db, float* partials,
                        int batch, const Conv0BnFuse& bn, cudaStream_t s, int channels_per_thread = 0);
void launch_bn_relu_pool(const __half* y, const float* scale, const float* shift, __half* out, int n,
                         int h, int w, int c, int pool, int flatten_chw, uint8_t* q_out, uint8_t* q_scale,
                         const BnFusedFwd& fused, cudaStream_t s);
int launch_bn_bwd_reduce(const __half* y, const __half* g, const float* scale, const float* shift,
                         const float* mean, const float* invstd, float* partials, double* accum /*[2][C] or null*/,
                         int n, int h, int w, int c, int pool, int g_flatten_chw, cudaStream_t s);
void launch_chw_to_hwc(const __half* in, __half* out, int n, int c, int p, cudaStream_t s);
void launch_bn_bwd_apply(const __half* y, const __half* g, const float* scale, const float* shift,
                         const float* mean, const float* invstd, const float* coef, __half* dy,
                         float* dbias /*[C] atomics or null*/, const BnFusedBwd& fused, int n, int h, int w, int c,
                         int pool, int g_flatten_chw, cudaStream_t s);
void launch_transpose_taps(const __half* w, __half* wt, int c_out, int taps, int c_in, cudaStream_t s);

}  // namespace ddpt
