// Weight-gradient GEMM for convolutions and linear layers on sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces cuDNN convolution wgrad / cuBLAS linear wgrad.
//
//   dW[m, tap, c] = sum over output pixels p:  dY[p, m] * X[p + tap, c]
//
// The reduction runs over pixels, which are the *strided* dimension of both NHWC operands, so both
// are fed to the tensor core MN-major straight from TMA boxes (no transposes are materialised):
//   A stage : dY brick  [K = brick pixels][M = 128 channels] as two 64-channel column blocks
//   B stage : X  brick shifted by the filter tap [K = brick pixels][N = BLOCK_N channels]
// A CTA owns one (128-row M tile, tap, BLOCK_N column tile) accumulator in TMEM and a contiguous
// range of pixel bricks (split-K over gridDim.z); partial results are combined with 128-bit
// vector reductions (red.global.add.v4.f32) into the fp32 gradient arena.
#include <cuda.h>
#include <stdexcept>
#include "../include/ddpt_nn.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"
#include "tensor_map.h"

namespace ddpt {
namespace {

using namespace umma;

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kMaxSmem = 227 * 1024;

struct WgradArgs {
  int tiles_w, tiles_h, tiles_n;     // pixel bricks per dimension
  int box_w, box_h, box_n, brick_rows;
  int taps_w, num_taps;
  int c_total;                       // GEMM N total (input channels)
  int stages, slot_bytes;            // pipeline depth and bytes of one 64-channel column block
  float* dw;
  int use_red;
  int fp16;                          // operands are fp16 instead of bf16
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradArgs args) {
  const uint32_t kIdesc = make_idesc_bf16(kBlockM, BN, true, true) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);
  constexpr int kABlocks = kBlockM / 64, kBBlocks = BN / 64;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = (kABlocks + kBBlocks) * args.slot_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + args.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + args.stages;
  uint64_t* accum_bar = empty_bar + args.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int tap = blockIdx.y % args.num_taps;
  const int c0 = (blockIdx.y / args.num_taps) * BN;
  const int tr = tap / args.taps_w, ts = tap - tr * args.taps_w;

  const int total_bricks = args.tiles_w * args.tiles_h * args.tiles_n;
  const int per_split = (total_bricks + gridDim.z - 1) / gridDim.z;
  const int brick_begin = blockIdx.z * per_split;
  const int num_bricks = max(min(total_bricks, brick_begin + per_split) - brick_begin, 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // operands come from the previous kernel of the stream: wait for it only now, after the setup
  pdl_trigger();

  if (warp == 0) {
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      const uint32_t tx_bytes = (kABlocks + kBBlocks) * args.brick_rows * 128;
      int ring_s = 0;
      uint32_t ring_phase = 0;
      // brick coordinates advance with counters: runtime divisions on this one in-order thread cost more than the
      // MMAs of a brick (the producer lane was the bottleneck of the GEMM kernels, profiles/ncu_fwd10_pair_v1.txt)
      int bw = brick_begin % args.tiles_w;
      int bh = (brick_begin / args.tiles_w) % args.tiles_h;
      int bn = brick_begin / (args.tiles_w * args.tiles_h);
      for (int i = 0; i < num_bricks; ++i) {
        const int s = ring_s;
        const uint32_t phase = ring_phase;
        if (++ring_s == args.stages) { ring_s = 0; ring_phase ^= 1; }
        mbar_wait(&empty_bar[s], phase ^ 1, 11);
        const int w0 = bw * args.box_w, h0 = bh * args.box_h, n0 = bn * args.box_n;
        if (++bw == args.tiles_w) { bw = 0; if (++bh == args.tiles_h) { bh = 0; ++bn; } }
        uint8_t* stage = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
#pragma unroll
        for (int j = 0; j < kABlocks; ++j)
          tma_load_4d(stage + j * args.slot_bytes, &tmap_dy, &full_bar[s], m0 + j * 64, w0, h0, n0);
#pragma unroll
        for (int j = 0; j < kBBlocks; ++j)
          tma_load_4d(stage + (kABlocks + j) * args.slot_bytes, &tmap_x, &full_bar[s], c0 + j * 64, w0 + ts,
                      h0 + tr, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      const int k_steps = args.brick_rows / kUmmaK;
      int ring_s = 0;
      uint32_t ring_phase = 0;
      for (int i = 0; i < num_bricks; ++i) {
        const int s = ring_s;
        const uint32_t phase = ring_phase;
        if (++ring_s == args.stages) { ring_s = 0; ring_phase ^= 1; }
        mbar_wait(&full_bar[s], phase, 12);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + kABlocks * args.slot_bytes;
        uint64_t adesc = make_smem_desc(a_addr, args.slot_bytes, 1024);
        uint64_t bdesc = make_smem_desc(b_addr, args.slot_bytes, 1024);
        for (int k = 0; k < k_steps; ++k) {               // 16 pixels further per step: one add per descriptor
          mma_bf16_ss(tmem_base, adesc, bdesc, kIdesc, (i | k) != 0 ? 1u : 0u);
          adesc = desc_advance(adesc, kUmmaK * 128);
          bdesc = desc_advance(bdesc, kUmmaK * 128);
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(accum_bar);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(accum_bar, 0, 13);
    tc_fence_after();
    float* out_row = args.dw + (static_cast<long long>(m0 + row) * args.num_taps + tap) * args.c_total + c0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (num_bricks > 0) {
        float* o = out_row + c * 32;
        if (args.use_red) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(o + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

template <int BN>
void launch_impl(const WgradProblem& p, cudaStream_t stream) {
  WgradArgs a{};
  a.fp16 = p.fp16;
  a.box_w = p.box_w; a.box_h = p.box_h; a.box_n = p.box_n;
  a.brick_rows = p.box_w * p.box_h * p.box_n;
  a.tiles_w = (p.out_w + p.box_w - 1) / p.box_w;
  a.tiles_h = (p.out_h + p.box_h - 1) / p.box_h;
  a.tiles_n = (p.out_n + p.box_n - 1) / p.box_n;
  a.taps_w = p.taps_w; a.num_taps = p.taps_h * p.taps_w;
  a.c_total = p.c_total; a.dw = p.dw;
  a.use_red = (p.split_k > 1 || p.accumulate) ? 1 : 0;
  if (a.brick_rows % kUmmaK != 0 || a.brick_rows > 128 || p.m_total % kBlockM != 0 || p.c_total % BN != 0)
    throw std::runtime_error("wgrad_gemm: unsupported shape (brick %16, brick<=128, M%128, C%BLOCK_N)");
  a.slot_bytes = ((a.brick_rows * 128 + 1023) / 1024) * 1024;
  const int stage_bytes = (kBlockM / 64 + BN / 64) * a.slot_bytes;
  a.stages = (kMaxSmem - 2048) / stage_bytes;
  if (a.stages > 8) a.stages = 8;
  if (a.stages < 2) throw std::runtime_error("wgrad_gemm: tile does not fit shared memory twice");
  const size_t smem_bytes = 1024 + (size_t)a.stages * stage_bytes + (2 * a.stages + 1) * 8 + 16;

  const uint64_t dy_dims[4] = {(uint64_t)p.m_total, (uint64_t)p.out_w, (uint64_t)p.out_h, (uint64_t)p.out_n};
  const uint64_t dy_strides[3] = {(uint64_t)p.m_total * 2, (uint64_t)p.m_total * p.out_w * 2,
                                  (uint64_t)p.m_total * p.out_w * p.out_h * 2};
  const uint64_t x_dims[4] = {(uint64_t)p.c_total, (uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)p.out_n};
  const uint64_t x_strides[3] = {(uint64_t)p.c_total * 2, (uint64_t)p.c_total * p.in_w * 2,
                                 (uint64_t)p.c_total * p.in_w * p.in_h * 2};
  const uint32_t box[4] = {64u, (uint32_t)p.box_w, (uint32_t)p.box_h, (uint32_t)p.box_n};
  CUtensorMap tmap_dy = make_tensor_map_bf16(p.dy, 4, dy_dims, dy_strides, box);
  CUtensorMap tmap_x = make_tensor_map_bf16(p.x, 4, x_dims, x_strides, box);

  auto kernel = wgrad_gemm_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured = true;
  }
  dim3 grid(p.m_total / kBlockM, a.num_taps * (p.c_total / BN), p.split_k > 1 ? p.split_k : 1);
  launch_pdl(kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, tmap_dy, tmap_x, a);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void launch_wgrad_gemm(const WgradProblem& p, int block_n, cudaStream_t stream) {
  switch (block_n) {
    case 64: return launch_impl<64>(p, stream);
    case 128: return launch_impl<128>(p, stream);
    case 256: return launch_impl<256>(p, stream);
  }
  throw std::runtime_error("wgrad_gemm: block_n must be 64, 128 or 256");
}

}  // namespace ddpt
