// Tensor-core GEMM / convolution kernels (tcgen05 + TMEM + TMA) -- host API.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ddpt {

// One description covers: 3x3 "valid" convolution forward, its data-gradient (transposed taps,
// zero-filled halo), and plain GEMM / linear layers (1x1 taps, W = rows, H = N = 1).
//
//   D[m, n] = sum_{tap, c}  A[pixel(m) + sign * tap, c] * B[n, tap, c]
//
// A is an NHWC bf16 tensor addressed through a 4-D TMA map (C, W, H, N); an M tile is a
// (box_n x box_h x box_w) brick of output pixels, so every filter tap of the implicit GEMM is the
// same TMA box shifted by (dh, dw) -- no im2col buffer and no gather.  B is a 2-D bf16 matrix:
//   b_mn_major == 0:  [n_total, taps * C]  (row = output column n, K contiguous)       -> forward / linear
//   b_mn_major == 1:  [C_k,     taps * n_total] (row = reduction channel, n contiguous) -> data gradient
struct ConvGemmProblem {
  // A operand (activations)
  const __nv_bfloat16* a;
  int a_c, a_w, a_h, a_n;                  // extents of the NHWC tensor (C innermost)
  // output pixel grid and its tiling into bricks (box_w * box_h * box_n <= 128 rows per tile)
  int out_w, out_h, out_n;
  int box_w, box_h, box_n;
  // filter taps: tap (r, s) reads A at (h + a_h_off + sign * r, w + a_w_off + sign * s)
  int taps_h, taps_w, tap_sign, a_h_off, a_w_off;
  // B operand
  const __nv_bfloat16* b;
  int n_total;                             // GEMM N (multiple of 64)
  int b_mn_major;
  long long b_ld;                          // leading dimension of B in elements
  // epilogue
  void* out;                               // bf16 or fp32, row-major [out_n*out_h*out_w, ld_out]
  long long ld_out;
  int out_fp32;                            // 1: fp32 output (atomicAdd when split_k > 1)
  const float* bias;                       // optional, per output column
  int relu;
  float* stat_partials;                    // optional [num_m_tiles][2][n_total]: per-tile column sum / sum of squares
  double* stat_accum;                      // optional [2][n_total] fp64 accumulators (zeroed by the caller): the epilogues add
                                           // their column sums / sums of squares with red.global.add.f64 instead of writing
                                           // partial rows, and the consumer pass derives the BatchNorm coefficients itself
                                           // (no *_finalize launch).  Unordered adds: use stat_partials for bit-stable runs.
  int split_k;                             // K splits (fp32 output only)
  int fp16;                                // A / B operands (and a 16-bit output) are fp16, not bf16: same layouts, other format
  long long split_slab_stride;             // > 0: split z stores its partial tile at out + z * stride (plain stores, no
                                           // atomics); splitk_finish folds the slabs in order (+ bf16 pack + BN statistics)
};

// block_n in {64, 128, 256}; returns the number of M tiles launched.
int launch_conv_gemm(const ConvGemmProblem& p, int block_n, cudaStream_t stream);
// Persistent variant (one CTA per SM, double-buffered TMEM accumulator); no split-K.  With stat_partials the
// statistics are accumulated per CTA and the return value is the number of row groups written (= grid size).
int launch_conv_gemm_persistent(const ConvGemmProblem& p, int block_n, cudaStream_t stream, int cta_pair = 0);
// Halo-reuse variant: the activation brick (+halo) is loaded once per channel chunk and every filter tap is a
// shifted descriptor view; requires box_w == 8 and box_h * box_n <= 16.  Same return convention as persistent.
int launch_conv_gemm_halo(const ConvGemmProblem& p, int block_n, cudaStream_t stream);
// 3x3 data gradient with 64 input channels (n_total == 64, MN-major weights, tap_sign == -1, C_out 64 or 128): the three
// filter columns are merged into the MMA's N dimension (128 x 192 instructions) and summed with a pixel shift in the
// epilogue; a_h_off == a_w_off is the forward convolution's padding.  See csrc/nn/conv_dgrad_kwmerge.cu.
int launch_conv_dgrad_kwmerge(const ConvGemmProblem& p, cudaStream_t stream);

// Fold `splits` fp32 slabs [rows, cols] (stride slab_stride elements) in split order into a bf16 matrix and, optionally,
// BatchNorm column statistics as partial rows [blocks][2][cols] (returns the number of rows written) -- the second half
// of a split-K linear layer whose output tile grid alone cannot fill the GPU (Toy_Net dense.1: 64 tiles on 148 SMs).
int launch_splitk_finish(const float* slabs, int splits, long long slab_stride, int rows, int cols,
                         __nv_bfloat16* out, float* stat_partials, int fp16, cudaStream_t stream);
int splitk_finish_max_rows();

// Weight gradient: dW[m, tap, n] += sum_pixels dY[pixel, m] * X[pixel + tap, n]  (both operands MN-major,
// reduction over the pixel bricks; fp32 accumulate into a zero-initialised or beta=1 output).
struct WgradProblem {
  const __nv_bfloat16* dy;                 // [N, Ho, Wo, M] bf16
  int m_total, out_w, out_h, out_n;        // M = output channels (multiple of 128 or 64)
  const __nv_bfloat16* x;                  // [N, H, W, C] bf16
  int c_total, in_w, in_h;                 // C = input channels (GEMM N, multiple of 64)
  int taps_h, taps_w;
  int box_w, box_h, box_n;                 // reduction brick (multiple of 16 pixels, <= 128)
  float* dw;                               // [M, taps_h, taps_w, C] fp32
  int split_k;                             // number of pixel-brick partitions (vector red.add when > 1)
  int accumulate;                          // 1: add into dw even when split_k == 1 (beta = 1)
  int fp16;                                // dy / x are fp16 instead of bf16
  float* slab;                             // multitap only: optional [split_k][M*taps*C] fp32 workspace; partials are
                                           // stored there and folded by a reduction kernel instead of red.add
};
void launch_wgrad_gemm(const WgradProblem& p, int block_n, cudaStream_t stream);
// Halo-reuse variant: one CTA accumulates a group of filter taps from a single dY + X(halo) brick load.
// box_w must be a multiple of 8; the output is always accumulated (red.add) into dw.
void launch_wgrad_gemm_multitap(const WgradProblem& p, int block_n, cudaStream_t stream);

}  // namespace ddpt
