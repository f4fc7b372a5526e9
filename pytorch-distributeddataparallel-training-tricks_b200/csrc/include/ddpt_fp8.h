// Block-scaled fp8 (MXFP8: e4m3 data, one UE8M0 scale per 32 elements along K) -- host API.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ddpt {

// Quantise rows of a bf16 (or fp32) matrix [rows, cols] along cols in blocks of 32:
//   scale_e = clamp(ceil(log2(amax / 448)) + 127), q = sat_e4m3(x * 2^-(scale_e - 127))
// data_out: uint8 [rows, cols] (e4m3 bit patterns), scale_out: uint8 [rows, cols / 32] (UE8M0).
void launch_quantize_mxfp8(const void* src, int src_is_fp32, uint8_t* data_out, uint8_t* scale_out, long long rows,
                           int cols, cudaStream_t stream);

// Forward implicit GEMM (3x3 'valid' convolution taps or a plain linear layer) on block-scaled fp8 operands:
//   D[m, n] = sum_{tap, c} A[pixel(m) + tap, c] * B[n, tap, c],  A/B e4m3 with per-32-channel UE8M0 scales.
struct Mxfp8GemmProblem {
  const uint8_t* a;            // [N, H, W, C] e4m3
  const uint8_t* a_scale;      // [N, H, W, C / 32] UE8M0
  int a_c, a_w, a_h, a_n;      // C % 128 == 0
  int out_w, out_h, out_n;
  int box_w, box_h, box_n;     // <= 128 rows
  int taps_h, taps_w;
  const uint8_t* b;            // [n_total, taps * C] e4m3 (K contiguous)
  const uint8_t* b_scale;      // [n_total, taps * C / 32]
  int n_total;                 // % 128 == 0
  void* out;                   // bf16 [pixels, ld_out]
  long long ld_out;
  const float* bias;
  float* stat_partials;        // same contract as the persistent bf16 kernel
  int relu;
};
// Returns the number of statistic row groups written (grid size) when stat_partials != nullptr, else M tiles.
int launch_conv_gemm_mxfp8(const Mxfp8GemmProblem& p, cudaStream_t stream);

}  // namespace ddpt
