// Data gradient of a 3x3 stride-1 convolution with 64 INPUT channels (Toy_Net conv.3; reference model.py:20) for
// sm_100a: tcgen05 + TMEM + TMA, the three filter COLUMNS merged into the MMA's N dimension.
//
// Why a dedicated kernel.  As an implicit GEMM the data gradient has N = C_in = 64.  A 128 x 64 x 16 tcgen05.mma is
// 32 tensor cycles but fetches 4 KB of A + 2 KB of B from shared memory, and the tensor core's shared-memory operand
// pipe moves 128 B/clk: 48 cycles per instruction, a 67 % ceiling (profiles/ncu_r2_dgrad3_halo.txt: tensor pipe 48.6 %,
// l1tex__data_pipe_tc_wavefronts_mem_shared 73 %, the issuing lane permanently back-pressured) -- and the activation
// brick is re-read for every filter tap.  Here one MMA covers the three taps of a filter ROW at once:
//
//     D[128 pixels, (kw, ci) = 192]  +=  dy[128 pixels, 16 co]  x  W[16 co, kh, (kw, ci)]
//
// i.e. the SAME activation view feeds three taps (10 KB of operands for 96 tensor cycles: compute bound again), and
// the activation brick + row halo is loaded once per 64-channel chunk instead of once per filter column.  Accumulator
// column block kw holds the contribution of dy pixel (y, x) to dx pixel (y + kh, x + kw): the three blocks are summed
// with a shift of kw pixels.  Tiles are 4 rows x 32 consecutive pixels, so the 32 lanes of an epilogue warp (= 32 TMEM
// lanes = one pixel row) are neighbours along x and the shift is two warp shuffles per value:
//
//     dx[y, x0 - 2 + L] = D0[L] + D1[L - 1] + D2[L - 2]            (lanes 2..31 -> 30 output columns per tile)
//
// The kh taps accumulate in TMEM through descriptor views shifted by whole pixel rows (32 x 128 B = 4 swizzle atoms:
// every view stays atom-aligned, see conv_gemm_halo.cu).  Out-of-image pixels are TMA zero fill, so a forward padding
// is just an origin offset.  Weights (9 x C_out x 64, <= 144 KB) stay resident in shared memory.
//
// Warp roles as in the other GEMM kernels: warp 0 TMA producer, warp 1 MMA issuer (one elected lane, fully unrolled),
// warps 2..9 epilogue on a double-buffered TMEM accumulator (2 x 192 columns).
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include "../include/ddpt_nn.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"
#include "tensor_map.h"

namespace ddpt {
namespace {

using namespace umma;

constexpr int kThreads = 320;
constexpr int kEpilogueWarps = 8;
constexpr int kTileRows = 4;                                 // dx rows per tile: 4 x 32 lanes = 128 MMA rows
constexpr int kTileCols = 30;                                // dx columns per tile (lanes 2..31)
constexpr int kBrickRows = kTileRows + 2;                    // + row halo of the 3 filter rows
constexpr int kRowBytes = 32 * 128;                          // one pixel row of a brick: 32 pixels x 64 channels
constexpr int kBrickBytes = kBrickRows * kRowBytes;          // 24 KB
constexpr int kBBlockBytes = 64 * 128;                       // 64 K rows (co) x 64 N columns (ci), MN-major
constexpr int kUmmaK = 16;
constexpr int kAccStride = 256;                              // TMEM columns between the two accumulators (192 used)
constexpr int kMaxSmem = 227 * 1024;

struct KwArgs {
  int out_w, out_h, out_n;          // dx extents
  int tiles_w, tiles_h;
  int pad;                          // padding of the forward convolution (0: 'valid')
  int a_stages;
  void* out;
  long long ld_out;
  int fp16;
};

template <int KC>                                            // C_out / 64 channel chunks (the GEMM's K per filter tap)
__global__ void __launch_bounds__(kThreads, 1)
conv_dgrad_kwmerge_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const KwArgs args, const int num_tiles) {
  constexpr int kBBytes = 9 * KC * kBBlockBytes;
  const uint32_t kIdesc = make_idesc_bf16(128, 192, false, true) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + args.a_stages * kBrickBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + kBBytes);
  uint64_t* a_empty = a_full + args.a_stages;
  uint64_t* tmem_full = a_empty + args.a_stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_ready = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < args.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpilogueWarps); }
    mbar_init(b_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int tiles_per_image = args.tiles_w * args.tiles_h;

  if (warp == 0) {
    // ===== TMA producer: the weights once, then one brick (4 rows + halo, 32 pixels, 64 channels) per (tile, chunk) =====
    if (elect_one()) {
      mbar_arrive_expect_tx(b_ready, kBBytes);
      for (int kh = 0; kh < 3; ++kh)
        for (int chunk = 0; chunk < KC; ++chunk)
          for (int kw = 0; kw < 3; ++kw)     // the three kw blocks of (kh, chunk) are consecutive: LBO = one block
            tma_load_2d(smem_b + ((kh * KC + chunk) * 3 + kw) * kBBlockBytes, &tmap_b, b_ready, (kh * 3 + kw) * 64,
                        chunk * 64);
      int a_s = 0;
      uint32_t a_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_image, rem = tile - n * tiles_per_image;
        const int th = rem / args.tiles_w, tw = rem - th * args.tiles_w;
        const int x_org = tw * kTileCols + args.pad - 2, y_org = th * kTileRows + args.pad - 2;
#pragma unroll
        for (int chunk = 0; chunk < KC; ++chunk) {
          mbar_wait(&a_empty[a_s], a_phase ^ 1, 61);
          mbar_arrive_expect_tx(&a_full[a_s], kBrickBytes);
          tma_load_4d(smem_a + a_s * kBrickBytes, &tmap_a, &a_full[a_s], chunk * 64, x_org, y_org, n);
          if (++a_s == args.a_stages) { a_s = 0; a_phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: 3 filter rows x 4 K steps per brick, each MMA 128 pixels x 192 (kw, ci) columns =====
    if (elect_one()) {
      int a_s = 0;
      uint32_t a_phase = 0, t = 0;
      mbar_wait(b_ready, 0, 62);
      const uint64_t bdesc_base = make_smem_desc(smem_u32(smem_b), kBBlockBytes, 1024);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 63);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
#pragma unroll
        for (int chunk = 0; chunk < KC; ++chunk) {
          mbar_wait(&a_full[a_s], a_phase, 64);
          tc_fence_after();
          const uint64_t adesc_base = make_smem_desc(smem_u32(smem_a + a_s * kBrickBytes), 0, 1024);
#pragma unroll
          for (int dr = 0; dr < 3; ++dr) {                   // brick row offset dr <-> filter row kh = 2 - dr
            const uint32_t b_off = static_cast<uint32_t>((((2 - dr) * KC + chunk) * 3) * kBBlockBytes);
            const uint64_t adesc_dr = desc_advance(adesc_base, dr * kRowBytes);
#pragma unroll
            for (int k = 0; k < 64 / kUmmaK; ++k)
              mma_bf16_ss(tmem_d, desc_advance(adesc_dr, k * (kUmmaK * 2)),
                          desc_advance(bdesc_base, b_off + k * (kUmmaK * 128)), kIdesc, (chunk | dr | k) != 0 ? 1u : 0u);
          }
          mma_commit(&a_empty[a_s]);
          if (++a_s == args.a_stages) { a_s = 0; a_phase ^= 1; }
        }
        mma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue: warp quadrant = tile row, lane = pixel, column half = 32 of the 64 input channels =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int n = tile / tiles_per_image, rem = tile - n * tiles_per_image;
      const int th = rem / args.tiles_w, tw = rem - th * args.tiles_w;
      const int y = th * kTileRows + quad, x = tw * kTileCols + lane - 2;
      const bool valid = lane >= 2 && x < args.out_w && y < args.out_h;
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase, 65);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(quad * 32) << 16) + half * 32;
      uint32_t v[32];
      float f[32];
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      tmem_ld_32x32(taddr + 64, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += __shfl_up_sync(0xffffffffu, __uint_as_float(v[j]), 1);
      tmem_ld_32x32(taddr + 128, v);
      tmem_ld_wait();
      // the accumulator is in registers: hand it back before the shuffles and stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += __shfl_up_sync(0xffffffffu, __uint_as_float(v[j]), 2);
      if (valid) {
        const long long out_row = (static_cast<long long>(n) * args.out_h + y) * args.out_w + x;
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(args.out) + out_row * args.ld_out + half * 32;
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

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int KC>
int launch_impl(const ConvGemmProblem& p, cudaStream_t stream) {
  KwArgs a{};
  a.out_w = p.out_w; a.out_h = p.out_h; a.out_n = p.out_n;
  a.tiles_w = (p.out_w + kTileCols - 1) / kTileCols;
  a.tiles_h = (p.out_h + kTileRows - 1) / kTileRows;
  a.pad = p.a_h_off;
  a.out = p.out; a.ld_out = p.ld_out; a.fp16 = p.fp16;
  constexpr int kBBytes = 9 * KC * kBBlockBytes;
  a.a_stages = std::min(6, (kMaxSmem - 2048 - kBBytes) / kBrickBytes);
  const size_t smem_bytes = 1024 + (size_t)a.a_stages * kBrickBytes + kBBytes + (2 * a.a_stages + 5) * 8 + 16;
  if (a.a_stages < 2 || smem_bytes > (size_t)kMaxSmem) throw std::runtime_error("conv_dgrad_kwmerge: shared memory budget exceeded");

  // dy as a (C, W, H, N) map, box (64 channels, 32 pixels, 4 + 2 rows, 1 image): rows of the brick are [row][pixel]
  const uint64_t a_dims[4] = {(uint64_t)p.a_c, (uint64_t)p.a_w, (uint64_t)p.a_h, (uint64_t)p.a_n};
  const uint64_t a_strides[3] = {(uint64_t)p.a_c * 2, (uint64_t)p.a_c * p.a_w * 2, (uint64_t)p.a_c * p.a_w * p.a_h * 2};
  const uint32_t a_box[4] = {64u, 32u, (uint32_t)kBrickRows, 1u};
  CUtensorMap tmap_a = make_tensor_map_bf16(p.a, 4, a_dims, a_strides, a_box);
  // weights [C_out, (kh, kw, ci)]: a 64 (ci) x 64 (co) box is one MN-major block of the B operand
  const uint64_t b_dims[2] = {(uint64_t)9 * 64, (uint64_t)p.a_c};
  const uint64_t b_strides[1] = {(uint64_t)p.b_ld * 2};
  const uint32_t b_box[2] = {64u, 64u};
  CUtensorMap tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);

  const int num_tiles = a.tiles_w * a.tiles_h * p.out_n;
  const int grid = std::min(num_tiles, sm_count());
  auto kernel = conv_dgrad_kwmerge_kernel<KC>;
  DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  launch_pdl(kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, tmap_a, tmap_b, a, num_tiles);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return num_tiles;
}

}  // namespace

int launch_conv_dgrad_kwmerge(const ConvGemmProblem& p, cudaStream_t stream) {
  if (!p.b_mn_major || p.taps_h != 3 || p.taps_w != 3 || p.tap_sign != -1 || p.n_total != 64 || p.out_fp32 ||
      p.bias != nullptr || p.relu || p.stat_partials != nullptr || p.stat_accum != nullptr || p.split_k > 1 ||
      p.a_h_off != p.a_w_off || p.a_h_off < 0 || p.a_h_off > 2)
    throw std::runtime_error("conv_dgrad_kwmerge: 3x3 data gradient with 64 input channels, MN-major weights, 16-bit output");
  switch (p.a_c) {
    case 64: return launch_impl<1>(p, stream);
    case 128: return launch_impl<2>(p, stream);
  }
  throw std::runtime_error("conv_dgrad_kwmerge: C_out must be 64 or 128 (resident weights)");
}

}  // namespace ddpt
