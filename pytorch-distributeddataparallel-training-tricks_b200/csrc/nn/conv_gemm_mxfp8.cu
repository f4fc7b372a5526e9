// Block-scaled fp8 (MXFP8) forward implicit GEMM for sm_100a: tcgen05.mma.kind::mxf8f6f4.block_scale.
//
// The "O1-fp8" AMP path (BASELINE.json config 3; SURVEY.md section 7.1 amp/): activations and weights are e4m3
// with one UE8M0 power-of-two scale per 32 elements along K; the tensor core applies both scales inside the
// MMA (K = 32 per instruction) and accumulates in fp32 TMEM.  Persistent kernel, same tiling as
// conv_gemm_persistent.cu (4-D TMA box per filter tap, double-buffered accumulator), with one extra role:
//
//   warp 0        TMA producer: fp8 activation brick (128 rows x 128 channels = 128 B rows) + weight tile
//   warp 1        MMA issuer: 4 x (K = 32) per K block, scale-factor byte k of the TMEM words for MMA k
//   warps 2..9    epilogue (bias / ReLU / BatchNorm partial statistics / bf16 store)
//   warps 10..13  scale-factor loaders: gather the K block's scale words (one 32-bit word = 4 x UE8M0 per row)
//                 from global memory and tcgen05.st them into TMEM, replicated in every 32-lane quadrant
//
// Scale-factor placement (measured with benchmarks/probe_umma_mx.py, profiles/umma_mxfp8_scale_layout_probe.txt):
// row m of A reads TMEM lane (m % 32) of its quadrant, column sfa + m / 32; column n of B reads lane (n % 32),
// column sfb + n / 32; byte k of the word belongs to the k-th K = 32 slice (sf_id field of the descriptor).
#include <cuda.h>
#include <stdexcept>
#include "../include/ddpt_fp8.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"
#include "tensor_map.h"

namespace ddpt {
namespace {

using namespace umma;

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kBlockKBytes = 128;           // 128 fp8 = one 128-byte swizzle row
constexpr int kUmmaKBytes = 32;             // K = 32 per MMA
constexpr int kStages = 6;
constexpr int kEpilogueWarps = 8;
constexpr int kSfWarps = 4;
constexpr int kThreads = 32 * (2 + kEpilogueWarps + kSfWarps);      // 448
constexpr int kAStageBytes = kBlockM * kBlockKBytes;                 // 16 KiB
constexpr int kBStageBytes = kBlockN * kBlockKBytes;                 // 16 KiB
constexpr int kSfColsPerStage = 4 + kBlockN / 32;                    // 4 (A) + 4 (B)
constexpr uint32_t kTmemCols = 512;                                  // 2 x 128 accumulators + 6 x 8 scale columns
constexpr int kSfBase = 2 * kBlockN;

struct FpArgs {
  int tiles_w, tiles_h;
  int box_w, box_h, box_n;
  int out_w, out_h, out_n;
  int in_w, in_h;
  int taps_w, num_taps;
  int k_chunks;                 // channels / 128
  int n_total;
  long long ld_out;
  __nv_bfloat16* out;
  const float* bias;
  float* stat_partials;
  const uint32_t* a_scale;      // [pixels][k_chunks] words
  const uint32_t* b_scale;      // [n_total][num_taps * k_chunks] words
  int relu, a_rows_bytes;
};

__device__ __forceinline__ uint32_t make_idesc_mxf8(uint32_t m, uint32_t n, uint32_t sf_id) {
  return ((sf_id & 3u) << 4) | ((n >> 3) << 17) | (1u << 23) | ((m >> 4) << 24) | ((sf_id & 3u) << 29);
}
__device__ __forceinline__ void mma_mxf8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const FpArgs args, const int num_m_tiles, const int num_n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kAStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * kBStageBytes);
  uint64_t* sf_full = full_bar + kStages;
  uint64_t* empty_bar = sf_full + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int total_kb = args.num_taps * args.k_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&sf_full[s], kSfWarps);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
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
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tile_m = tile / num_n_tiles, tile_n = tile - tile_m * num_n_tiles;
        const int tw = tile_m % args.tiles_w;
        const int th = (tile_m / args.tiles_w) % args.tiles_h;
        const int tn = tile_m / (args.tiles_w * args.tiles_h);
        const int w0 = tw * args.box_w, h0 = th * args.box_h, n0 = tn * args.box_n;
        int chunk = 0, sx = 0, r = 0;          // (filter tap, channel chunk) of the K block: counters, no divisions
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1, 61);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          tma_load_4d(smem_a + s * kAStageBytes, &tmap_a, &full_bar[s], chunk * kBlockKBytes, w0 + sx, h0 + r, n0);
          tma_load_2d(smem_b + s * kBStageBytes, &tmap_b, &full_bar[s], kb * kBlockKBytes, tile_n * kBlockN);
          if (++chunk == args.k_chunks) { chunk = 0; if (++sx == args.taps_w) { sx = 0; ++r; } }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 64);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t phase = (it / kStages) & 1;
          mbar_wait(&full_bar[s], phase, 62);
          mbar_wait(&sf_full[s], phase, 63);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * kAStageBytes);
          const uint32_t b_addr = smem_u32(smem_b + s * kBStageBytes);
          const uint32_t sfa = tmem_base + kSfBase + s * kSfColsPerStage;
#pragma unroll
          for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k) {
            const uint64_t adesc = make_smem_desc(a_addr + k * kUmmaKBytes, 0, 1024);
            const uint64_t bdesc = make_smem_desc(b_addr + k * kUmmaKBytes, 0, 1024);
            mma_mxf8_ss(tmem_d, adesc, bdesc, make_idesc_mxf8(kBlockM, kBlockN, k), sfa, sfa + 4, (kb | k) != 0 ? 1u : 0u);
          }
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp >= 2 + kEpilogueWarps) {
    // ===== scale-factor loaders: one warp per TMEM lane quadrant, identical content in every quadrant =====
    const int quad = warp & 3;
    const int box_hw = args.box_h * args.box_w;
    const uint32_t lane_bits = static_cast<uint32_t>(quad * 32) << 16;
    const int b_words_per_row = args.num_taps * args.k_chunks;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tile_m = tile / num_n_tiles, tile_n = tile - tile_m * num_n_tiles;
      const int tw = tile_m % args.tiles_w;
      const int th = (tile_m / args.tiles_w) % args.tiles_h;
      const int tn = tile_m / (args.tiles_w * args.tiles_h);
      // the four accumulator rows this lane describes: c * 32 + lane
      long long pix[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int row = c * 32 + lane;
        const int dn = row / box_hw, rem = row - dn * box_hw;
        const int dh = rem / args.box_w, dw = rem - dh * args.box_w;
        const int n = tn * args.box_n + dn, h = th * args.box_h + dh, w = tw * args.box_w + dw;
        const bool ok = dn < args.box_n && n < args.out_n && h < args.out_h && w < args.out_w;
        pix[c] = ok ? (static_cast<long long>(n) * args.in_h + h) * args.in_w + w : -1;
      }
      const uint32_t* b_rows[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        b_rows[c] = args.b_scale + static_cast<long long>(tile_n * kBlockN + c * 32 + lane) * b_words_per_row;
      int chunk_c = 0, sx_c = 0, r_c = 0;
      for (int kb = 0; kb < total_kb; ++kb, ++it) {
        const int s = it % kStages;
        const int chunk = chunk_c, sx = sx_c, r = r_c;
        if (++chunk_c == args.k_chunks) { chunk_c = 0; if (++sx_c == args.taps_w) { sx_c = 0; ++r_c; } }
        uint32_t wa[4], wb[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          wa[c] = pix[c] >= 0 ? __ldg(args.a_scale + (pix[c] + static_cast<long long>(r) * args.in_w + sx) * args.k_chunks + chunk)
                              : 0x7F7F7F7Fu;
          wb[c] = __ldg(b_rows[c] + kb);
        }
        mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1, 65);       // MMAs that read this slot have retired
        tc_fence_after();
        const uint32_t col = tmem_base + lane_bits + kSfBase + s * kSfColsPerStage;
        tmem_st_x4(col, wa[0], wa[1], wa[2], wa[3]);
        tmem_st_x4(col + 4, wb[0], wb[1], wb[2], wb[3]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sf_full[s]);
      }
    }
  } else {
    // ===== epilogue (identical to the bf16 persistent kernel, BN = 128) =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int box_hw = args.box_h * args.box_w;
    const int dn = row / box_hw, rem = row - dn * box_hw;
    const int dh = rem / args.box_w, dw = rem - dh * args.box_w;
    constexpr int kChunksPerWarp = (kBlockN / 32) / 2;
    float acc_s1[kChunksPerWarp], acc_s2[kChunksPerWarp];
#pragma unroll
    for (int cc = 0; cc < kChunksPerWarp; ++cc) acc_s1[cc] = acc_s2[cc] = 0.f;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int tile_m = tile / num_n_tiles, tile_n = tile - tile_m * num_n_tiles;
      const int tw = tile_m % args.tiles_w;
      const int th = (tile_m / args.tiles_w) % args.tiles_h;
      const int tn = tile_m / (args.tiles_w * args.tiles_h);
      const int n = tn * args.box_n + dn, h = th * args.box_h + dh, w = tw * args.box_w + dw;
      const bool valid = dn < args.box_n && n < args.out_n && h < args.out_h && w < args.out_w;
      const long long out_row = (static_cast<long long>(n) * args.out_h + h) * args.out_w + w;
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase, 66);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int c = half * kChunksPerWarp + cc;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * kBlockN + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const int col = tile_n * kBlockN + c * 32;
        if (args.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(args.bias + col + j);
        }
        if (args.stat_partials != nullptr) {
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
              float keep1 = upper ? s1[j + width] : s1[j];
              float give1 = upper ? s1[j] : s1[j + width];
              float keep2 = upper ? s2[j + width] : s2[j];
              float give2 = upper ? s2[j] : s2[j + width];
              s1[j] = keep1 + __shfl_xor_sync(0xffffffffu, give1, width);
              s2[j] = keep2 + __shfl_xor_sync(0xffffffffu, give2, width);
            }
          }
          acc_s1[cc] += s1[0];
          acc_s2[cc] += s2[0];
        }
        if (args.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (valid) {
          __nv_bfloat16* o = args.out + out_row * args.ld_out + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pk.x = pack_bf16x2(f[j], f[j + 1]);
            pk.y = pack_bf16x2(f[j + 2], f[j + 3]);
            pk.z = pack_bf16x2(f[j + 4], f[j + 5]);
            pk.w = pack_bf16x2(f[j + 6], f[j + 7]);
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (args.stat_partials != nullptr) {
      const int my_tile_n = blockIdx.x % num_n_tiles;
      float* dst = args.stat_partials + (static_cast<long long>(blockIdx.x) * 4 + quad) * 2 * args.n_total;
      for (int tn2 = 0; tn2 < num_n_tiles; ++tn2) {
#pragma unroll
        for (int cc = 0; cc < kChunksPerWarp; ++cc) {
          const int col = tn2 * kBlockN + (half * kChunksPerWarp + cc) * 32 + lane;
          dst[col] = tn2 == my_tile_n ? acc_s1[cc] : 0.f;
          dst[args.n_total + col] = tn2 == my_tile_n ? acc_s2[cc] : 0.f;
        }
      }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

int launch_conv_gemm_mxfp8(const Mxfp8GemmProblem& p, cudaStream_t stream) {
  if (p.a_c % kBlockKBytes != 0 || p.n_total % kBlockN != 0 || p.box_w * p.box_h * p.box_n > kBlockM)
    throw std::runtime_error("conv_gemm_mxfp8: needs C % 128 == 0, N % 128 == 0, brick <= 128 rows");
  FpArgs a{};
  a.box_w = p.box_w; a.box_h = p.box_h; a.box_n = p.box_n;
  a.out_w = p.out_w; a.out_h = p.out_h; a.out_n = p.out_n;
  a.in_w = p.a_w; a.in_h = p.a_h;
  a.tiles_w = (p.out_w + p.box_w - 1) / p.box_w;
  a.tiles_h = (p.out_h + p.box_h - 1) / p.box_h;
  const int tiles_n = (p.out_n + p.box_n - 1) / p.box_n;
  a.taps_w = p.taps_w; a.num_taps = p.taps_h * p.taps_w;
  a.k_chunks = p.a_c / kBlockKBytes;
  a.n_total = p.n_total; a.ld_out = p.ld_out; a.out = static_cast<__nv_bfloat16*>(p.out);
  a.bias = p.bias; a.stat_partials = p.stat_partials; a.relu = p.relu;
  a.a_scale = reinterpret_cast<const uint32_t*>(p.a_scale);
  a.b_scale = reinterpret_cast<const uint32_t*>(p.b_scale);
  a.a_rows_bytes = p.box_w * p.box_h * p.box_n * kBlockKBytes;

  const uint64_t a_dims[4] = {(uint64_t)p.a_c, (uint64_t)p.a_w, (uint64_t)p.a_h, (uint64_t)p.a_n};
  const uint64_t a_strides[3] = {(uint64_t)p.a_c, (uint64_t)p.a_c * p.a_w, (uint64_t)p.a_c * p.a_w * p.a_h};
  const uint32_t a_box[4] = {(uint32_t)kBlockKBytes, (uint32_t)p.box_w, (uint32_t)p.box_h, (uint32_t)p.box_n};
  CUtensorMap tmap_a = make_tensor_map_u8(p.a, 4, a_dims, a_strides, a_box);
  const uint64_t b_dims[2] = {(uint64_t)a.num_taps * p.a_c, (uint64_t)p.n_total};
  const uint64_t b_strides[1] = {(uint64_t)a.num_taps * p.a_c};
  const uint32_t b_box[2] = {(uint32_t)kBlockKBytes, (uint32_t)kBlockN};
  CUtensorMap tmap_b = make_tensor_map_u8(p.b, 2, b_dims, b_strides, b_box);

  const size_t smem_bytes = 1024 + kStages * (kAStageBytes + kBStageBytes) + (3 * kStages + 4) * 8 + 16;
  static bool configured = false;
  if (!configured) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    configured = true;
  }
  const int tiles_m = a.tiles_w * a.tiles_h * tiles_n;
  const int tiles_n_cols = p.n_total / kBlockN;
  const int total_tiles = tiles_m * tiles_n_cols;
  int grid = total_tiles < sm_count() ? total_tiles : sm_count();
  grid -= grid % tiles_n_cols;
  if (grid < tiles_n_cols) grid = tiles_n_cols;
  launch_pdl(conv_gemm_mxfp8_kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, tmap_a, tmap_b, a, tiles_m, tiles_n_cols);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return p.stat_partials != nullptr ? grid : tiles_m;
}

}  // namespace ddpt
