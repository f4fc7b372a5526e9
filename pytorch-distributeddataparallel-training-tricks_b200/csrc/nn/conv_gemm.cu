// Implicit-GEMM convolution / linear forward and data-gradient for sm_100a.
//
// Replaces cuDNN convolution fwd/dgrad and cuBLAS linear fwd/dgrad of the reference stack
// (Toy_Net layers K3/K6/K8/K11/K13).  One CTA computes a 128 x BLOCK_N output
// tile with the accumulator in TMEM:
//
//   warp 0      TMA producer: per K block (one filter tap x 64 input channels) one 4-D box load of
//               the activation brick shifted by the tap, plus the matching weight tile
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per K block),
//               tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld 32 columns at a time, + bias, ReLU, BatchNorm partial
//               statistics (column sum / sum of squares per tile), bf16 or fp32 store
//
// Operands are staged in 128-byte-swizzled shared memory exactly as TMA writes them; the matrix
// descriptors describe that layout (K-major for activations, K- or MN-major for weights).
#include <cuda.h>
#include "../include/ddpt_nn.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"
#include "tensor_map.h"
#include <stdexcept>

namespace ddpt {
namespace {

using namespace umma;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;               // producer, mma, 4 epilogue warps
constexpr int kAStageBytes = kBlockM * kBlockK * 2;   // 16 KiB

template <int BN> __host__ __device__ constexpr int stages_for() { return BN == 256 ? 4 : (BN == 128 ? 6 : 8); }

struct KernelArgs {
  int tiles_w, tiles_h;
  int box_w, box_h, box_n;
  int out_w, out_h, out_n;
  int taps_w, num_taps, tap_sign, a_h_off, a_w_off;
  int k_chunks;                 // reduction channels / 64
  int n_total;
  long long ld_out;
  void* out;
  const float* bias;
  float* stat_partials;
  double* stat_accum;
  int relu, out_fp32, a_rows_bytes;
  long long slab_stride;        // split-K: > 0 = every split stores its own slab (no atomics)
  int fp16;                     // operands and 16-bit outputs are fp16 instead of bf16
};

template <int BN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const KernelArgs args) {
  constexpr int kStages = stages_for<BN>();
  constexpr int kBStageBytes = BN * kBlockK * 2;
  const uint32_t kIdesc = make_idesc_bf16(kBlockM, BN, false, B_MN) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kAStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * kBStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates -------------------------------------------------------------------------
  const int tile_m = blockIdx.x;
  const int tw = tile_m % args.tiles_w;
  const int th = (tile_m / args.tiles_w) % args.tiles_h;
  const int tn = tile_m / (args.tiles_w * args.tiles_h);
  const int w0 = tw * args.box_w, h0 = th * args.box_h, n0 = tn * args.box_n;
  const int col0 = blockIdx.y * BN;

  const int total_kb = args.num_taps * args.k_chunks;
  const int kb_per_split = (total_kb + gridDim.z - 1) / gridDim.z;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(total_kb, kb_begin + kb_per_split);
  const int num_kb = max(kb_end - kb_begin, 0);

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // operands come from the previous kernel of the stream: wait for it only now, after the setup
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      const uint32_t tx_bytes = args.a_rows_bytes + kBStageBytes;
      // (tap, channel chunk) of the K block advance with counters: divisions once, before the loop
      int tap = kb_begin / args.k_chunks, chunk = kb_begin - tap * args.k_chunks;
      int r = tap / args.taps_w, sx = tap - r * args.taps_w;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kStages;
        const uint32_t phase = (i / kStages) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1, 1);
        const int kb = kb_begin + i;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        tma_load_4d(smem_a + s * kAStageBytes, &tmap_a, &full_bar[s], chunk * kBlockK,
                    w0 + args.a_w_off + args.tap_sign * sx, h0 + args.a_h_off + args.tap_sign * r, n0);
        if (!B_MN) {
          tma_load_2d(smem_b + s * kBStageBytes, &tmap_b, &full_bar[s], kb * kBlockK, col0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(smem_b + s * kBStageBytes + j * (kBlockK * 128), &tmap_b, &full_bar[s],
                        tap * args.n_total + col0 + j * 64, chunk * kBlockK);
        }
        if (++chunk == args.k_chunks) { chunk = 0; ++tap; if (++sx == args.taps_w) { sx = 0; ++r; } }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kStages;
        const uint32_t phase = (i / kStages) & 1;
        mbar_wait(&full_bar[s], phase, 2);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + s * kAStageBytes);
        const uint32_t b_addr = smem_u32(smem_b + s * kBStageBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint64_t adesc = make_smem_desc(a_addr + k * (kUmmaK * 2), 0, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc(b_addr + k * (kUmmaK * 128), kBlockK * 128, 1024)
                                      : make_smem_desc(b_addr + k * (kUmmaK * 2), 0, 1024);
          mma_bf16_ss(tmem_base, adesc, bdesc, kIdesc, (i | k) != 0 ? 1u : 0u);
        }
        mma_commit(&empty_bar[s]);          // smem stage reusable once these MMAs retire
      }
      mma_commit(accum_bar);                // accumulator complete
    }
  } else {
    // ===== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                       // accumulator row == TMEM lane
    const int box_hw = args.box_h * args.box_w;
    const int dn = row / box_hw, rem = row - dn * box_hw;
    const int dh = rem / args.box_w, dw = rem - dh * args.box_w;
    const int n = n0 + dn, h = h0 + dh, w = w0 + dw;
    const bool valid = num_kb > 0 && dn < args.box_n && n < args.out_n && h < args.out_h && w < args.out_w;
    const long long out_row = (static_cast<long long>(n) * args.out_h + h) * args.out_w + w;

    mbar_wait(accum_bar, 0, 3);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const int col = col0 + c * 32;
      if (args.bias != nullptr && blockIdx.z == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += __ldg(args.bias + col + j);
      }
      if (args.stat_partials != nullptr || args.stat_accum != nullptr) {
        // column statistics over the 32 rows of this warp: transpose-reduce (31 shuffles per quantity)
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s1[j] = valid ? f[j] : 0.f;
          s2[j] = s1[j] * s1[j];
        }
#pragma unroll
        for (int width = 16; width >= 1; width >>= 1) {
          const bool upper = (lane & width) != 0;
#pragma unroll
          for (int j = 0; j < width; ++j) {
            // keep half of the columns, hand the other half to the partner lane
            float keep1 = upper ? s1[j + width] : s1[j];
            float give1 = upper ? s1[j] : s1[j + width];
            float keep2 = upper ? s2[j + width] : s2[j];
            float give2 = upper ? s2[j] : s2[j + width];
            s1[j] = keep1 + __shfl_xor_sync(0xffffffffu, give1, width);
            s2[j] = keep2 + __shfl_xor_sync(0xffffffffu, give2, width);
          }
        }
        // lane L now holds the totals of column bitreverse-free index: column = lane (see note below)
        if (args.stat_accum != nullptr) {
          atomicAdd(args.stat_accum + col + lane, (double)s1[0]);
          atomicAdd(args.stat_accum + args.n_total + col + lane, (double)s2[0]);
        } else {
          float* dst = args.stat_partials + (static_cast<long long>(tile_m) * 4 + quad) * 2 * args.n_total;
          dst[col + lane] = s1[0];
          dst[args.n_total + col + lane] = s2[0];
        }
      }
      if (args.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (valid) {
        if (args.out_fp32) {
          float* o = static_cast<float*>(args.out) + out_row * args.ld_out + col;
          if (args.slab_stride > 0) o += static_cast<long long>(blockIdx.z) * args.slab_stride;
          if (gridDim.z > 1 && args.slab_stride == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(o + j, f[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
        } else {
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(args.out) + out_row * args.ld_out + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            if (args.fp16) {
              pk.x = pack_f16x2(f[j], f[j + 1]); pk.y = pack_f16x2(f[j + 2], f[j + 3]);
              pk.z = pack_f16x2(f[j + 4], f[j + 5]); pk.w = pack_f16x2(f[j + 6], f[j + 7]);
            } else {
              pk.x = pack_bf16x2(f[j], f[j + 1]); pk.y = pack_bf16x2(f[j + 2], f[j + 3]);
              pk.z = pack_bf16x2(f[j + 4], f[j + 5]); pk.w = pack_bf16x2(f[j + 6], f[j + 7]);
            }
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
  }
}

template <int BN, bool B_MN>
int launch_impl(const ConvGemmProblem& p, cudaStream_t stream) {
  constexpr int kStages = stages_for<BN>();
  constexpr int kBStageBytes = BN * kBlockK * 2;
  const size_t smem_bytes = 1024 + kStages * (kAStageBytes + kBStageBytes) + (2 * kStages + 1) * 8 + 16;

  KernelArgs a{};
  a.box_w = p.box_w; a.box_h = p.box_h; a.box_n = p.box_n;
  a.out_w = p.out_w; a.out_h = p.out_h; a.out_n = p.out_n;
  a.tiles_w = (p.out_w + p.box_w - 1) / p.box_w;
  a.tiles_h = (p.out_h + p.box_h - 1) / p.box_h;
  const int tiles_n = (p.out_n + p.box_n - 1) / p.box_n;
  a.taps_w = p.taps_w; a.num_taps = p.taps_h * p.taps_w; a.tap_sign = p.tap_sign;
  a.a_h_off = p.a_h_off; a.a_w_off = p.a_w_off;
  a.k_chunks = p.a_c / kBlockK;
  a.n_total = p.n_total; a.ld_out = p.ld_out; a.out = p.out; a.bias = p.bias;
  a.stat_partials = p.stat_partials; a.stat_accum = p.stat_accum; a.relu = p.relu; a.out_fp32 = p.out_fp32;
  a.a_rows_bytes = p.box_w * p.box_h * p.box_n * kBlockK * 2;
  a.slab_stride = p.split_k > 1 ? p.split_slab_stride : 0;
  a.fp16 = p.fp16;

  if (p.a_c % kBlockK != 0 || p.n_total % BN != 0 || p.box_w * p.box_h * p.box_n > kBlockM)
    throw std::runtime_error("conv_gemm: unsupported shape (C % 64, N % BLOCK_N, brick <= 128 rows)");
  if (p.split_k > 1 && !p.out_fp32)
    throw std::runtime_error("conv_gemm: split_k needs fp32 output");

  // A: (C, W, H, N) bf16, box (64, box_w, box_h, box_n), 128B swizzle, OOB -> 0
  const uint64_t a_dims[4] = {(uint64_t)p.a_c, (uint64_t)p.a_w, (uint64_t)p.a_h, (uint64_t)p.a_n};
  const uint64_t a_strides[3] = {(uint64_t)p.a_c * 2, (uint64_t)p.a_c * p.a_w * 2,
                                 (uint64_t)p.a_c * p.a_w * p.a_h * 2};
  const uint32_t a_box[4] = {(uint32_t)kBlockK, (uint32_t)p.box_w, (uint32_t)p.box_h, (uint32_t)p.box_n};
  CUtensorMap tmap_a = make_tensor_map_bf16(p.a, 4, a_dims, a_strides, a_box);

  CUtensorMap tmap_b;
  if (!B_MN) {
    const uint64_t b_dims[2] = {(uint64_t)a.num_taps * p.a_c, (uint64_t)p.n_total};
    const uint64_t b_strides[1] = {(uint64_t)p.b_ld * 2};
    const uint32_t b_box[2] = {(uint32_t)kBlockK, (uint32_t)BN};
    tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);
  } else {
    const uint64_t b_dims[2] = {(uint64_t)a.num_taps * p.n_total, (uint64_t)p.a_c};
    const uint64_t b_strides[1] = {(uint64_t)p.b_ld * 2};
    const uint32_t b_box[2] = {64u, (uint32_t)kBlockK};
    tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);
  }

  auto kernel = conv_gemm_kernel<BN, B_MN>;
  static bool configured = false;
  if (!configured) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    configured = true;
  }
  const int tiles_m = a.tiles_w * a.tiles_h * tiles_n;
  dim3 grid(tiles_m, p.n_total / BN, p.split_k > 1 ? p.split_k : 1);
  launch_pdl(kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, tmap_a, tmap_b, a);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return tiles_m;
}

}  // namespace

int launch_conv_gemm(const ConvGemmProblem& p, int block_n, cudaStream_t stream) {
  if (p.b_mn_major) {
    switch (block_n) {
      case 64: return launch_impl<64, true>(p, stream);
      case 128: return launch_impl<128, true>(p, stream);
      case 256: return launch_impl<256, true>(p, stream);
    }
  } else {
    switch (block_n) {
      case 64: return launch_impl<64, false>(p, stream);
      case 128: return launch_impl<128, false>(p, stream);
      case 256: return launch_impl<256, false>(p, stream);
    }
  }
  throw std::runtime_error("conv_gemm: block_n must be 64, 128 or 256");
}


// ---- split-K second half ------------------------------------------------------------------------------------------
namespace {
constexpr int kFinishThreads = 256;

__global__ void __launch_bounds__(kFinishThreads)
splitk_finish_kernel(const float* __restrict__ slabs, int splits, long long slab_stride, int rows, int cols,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ stat_partials, int fp16) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_part[];                       // [rows_per_blk][2][cols] ordered block fold
  const int cv = cols >> 3;
  const int v = threadIdx.x % cv, lane_row = threadIdx.x / cv, rows_per_blk = kFinishThreads / cv;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (int r = blockIdx.x * rows_per_blk + lane_row; r < rows; r += gridDim.x * rows_per_blk) {
    const float* src = slabs + (long long)r * cols + v * 8;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int z = 0; z < splits; ++z) {                    // fixed split order: deterministic
      const float4 x = __ldg(reinterpret_cast<const float4*>(src + z * slab_stride));
      const float4 y = __ldg(reinterpret_cast<const float4*>(src + z * slab_stride) + 1);
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
    }
    uint4 pk;
    if (fp16) { pk.x = pack_f16x2(a.x, a.y); pk.y = pack_f16x2(a.z, a.w); pk.z = pack_f16x2(b.x, b.y); pk.w = pack_f16x2(b.z, b.w); }
    else { pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w); pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w); }
    *reinterpret_cast<uint4*>(out + (long long)r * cols + v * 8) = pk;
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
  }
  if (stat_partials == nullptr) return;
  float* mine = s_part + (size_t)lane_row * 2 * cols;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mine[v * 8 + j] = s1[j]; mine[cols + v * 8 + j] = s2[j]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * cols; i += kFinishThreads) {
    float tot = 0.f;
    for (int r = 0; r < rows_per_blk; ++r) tot += s_part[(size_t)r * 2 * cols + i];
    stat_partials[(size_t)blockIdx.x * 2 * cols + i] = tot;
  }
}
}  // namespace

int splitk_finish_max_rows() { return sm_count() * 2; }

int launch_splitk_finish(const float* slabs, int splits, long long slab_stride, int rows, int cols,
                         __nv_bfloat16* out, float* stat_partials, int fp16, cudaStream_t stream) {
  if (cols % 8 != 0 || kFinishThreads % (cols / 8) != 0) throw std::runtime_error("splitk_finish: cols must be 8 x a divisor of 256");
  const int rows_per_blk = kFinishThreads / (cols / 8);
  int grid = (rows + rows_per_blk - 1) / rows_per_blk;
  if (grid > splitk_finish_max_rows()) grid = splitk_finish_max_rows();
  const size_t smem = (size_t)rows_per_blk * 2 * cols * sizeof(float);
  launch_pdl(splitk_finish_kernel, dim3(grid), dim3(kFinishThreads), smem, stream, slabs, splits, slab_stride, rows, cols, out,
             stat_partials, fp16);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return grid;
}

}  // namespace ddpt
