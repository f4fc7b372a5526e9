// Multi-tap weight-gradient GEMM for sm_100a (tcgen05 + TMEM + TMA): halo reuse of the activation brick.
//
//   dW[m, tap, c] = sum over output pixels p:  dY[p, m] * X[p + tap, c]
//
// wgrad_gemm.cu gives every filter tap its own CTA, so the dY brick is fetched 9 times and a shifted X brick
// once per tap -- the kernel is L2-bandwidth bound (24 KiB per 128 MMA cycles for C_in = 64).  Here one CTA
// owns a *group* of taps (as many as fit TMEM: taps x BLOCK_N <= 512 columns).  Per reduction brick it loads
//   - the dY brick once                       [K = brick pixels][128 channels]   (MN-major A)
//   - the X brick once, including its halo    [(bh + nr - 1) x bn x (bw + ns - 1) pixels][BLOCK_N channels]
// and issues one MMA chain per tap whose B descriptor is a *shifted view* of the halo tile: the start address
// moves by whole 128-byte rows and the 8-row K groups are (bw + ns - 1) * 128 bytes apart.  That is legal
// because the tensor core applies the 128B swizzle to absolute shared-memory addresses, exactly like TMA
// (measured: profiles/umma_descriptor_probe.txt), so no copy or re-layout is needed.
// L2 traffic per MMA cycle drops ~5x for Toy_Net's conv.3 and the kernel becomes tensor-core bound.
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
constexpr int kMaxGroups = 9;

struct TapGroup {
  int r0, nr, s0, ns;          // halo box: filter rows [r0, r0+nr) x columns [s0, s0+ns)
  int mask;                    // which of the nr*ns taps inside the box this group accumulates (bit = dr*ns + ds)
  int map;                     // which X tensor map (box shape) this group uses
  int halo_rows;               // (bh + nr - 1) * bn * (bw + ns - 1)
};

struct MtArgs {
  int tiles_w, tiles_h, tiles_n;
  int box_w, box_h, box_n, brick_rows;
  int taps_w, num_taps;
  int c_total;
  int stages, slot_a, slot_x;
  int num_groups;
  TapGroup groups[kMaxGroups];
  float* dw;                   // final output (red.add) when slab == nullptr
  float* slab;                 // optional [gridDim.z][M * taps * C] partial slabs written with plain stores
  long long slab_elems;
  int plain;                   // no slab, single split, beta = 0: plain stores straight into dw
  int fp16;                    // operands are fp16 instead of bf16
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int TMEM_COLS>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_multitap_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x0,
                      const __grid_constant__ CUtensorMap tmap_x1, const MtArgs args) {
  const uint32_t kIdesc = make_idesc_bf16(kBlockM, BN, true, true) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);
  constexpr int kABlocks = kBlockM / 64, kBBlocks = BN / 64;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kABlocks * args.slot_a + kBBlocks * args.slot_x;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + args.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + args.stages;
  uint64_t* accum_bar = empty_bar + args.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int gidx = blockIdx.y % args.num_groups;
  const int c0 = (blockIdx.y / args.num_groups) * BN;
  const TapGroup grp = args.groups[gidx];
  const int box_taps = grp.nr * grp.ns;

  const int total_bricks = args.tiles_w * args.tiles_h * args.tiles_n;
  const int per_split = (total_bricks + gridDim.z - 1) / gridDim.z;
  const int brick_begin = blockIdx.z * per_split;
  const int num_bricks = max(min(total_bricks, brick_begin + per_split) - brick_begin, 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(grp.map == 0 ? &tmap_x0 : &tmap_x1);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // operands come from the previous kernel of the stream: wait for it only now, after the setup
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer: dY brick + X halo brick per reduction step =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      const CUtensorMap* tmx = grp.map == 0 ? &tmap_x0 : &tmap_x1;
      const uint32_t tx_bytes = (kABlocks * args.brick_rows + kBBlocks * grp.halo_rows) * 128;
      int ring_s = 0;
      uint32_t ring_phase = 0;
      // brick coordinates advance with counters (no runtime divisions on the single producer lane)
      int bw = brick_begin % args.tiles_w;
      int bh = (brick_begin / args.tiles_w) % args.tiles_h;
      int bn = brick_begin / (args.tiles_w * args.tiles_h);
      for (int i = 0; i < num_bricks; ++i) {
        const int s = ring_s;
        const uint32_t phase = ring_phase;
        if (++ring_s == args.stages) { ring_s = 0; ring_phase ^= 1; }
        mbar_wait(&empty_bar[s], phase ^ 1, 31);
        const int w0 = bw * args.box_w, h0 = bh * args.box_h, n0 = bn * args.box_n;
        if (++bw == args.tiles_w) { bw = 0; if (++bh == args.tiles_h) { bh = 0; ++bn; } }
        uint8_t* stage = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        // tensor maps are declared (C, W, N, H): the brick's pixel order in shared memory is w, then n, then h
#pragma unroll
        for (int j = 0; j < kABlocks; ++j)
          tma_load_4d(stage + j * args.slot_a, &tmap_dy, &full_bar[s], m0 + j * 64, w0, n0, h0);
#pragma unroll
        for (int j = 0; j < kBBlocks; ++j)
          tma_load_4d(stage + kABlocks * args.slot_a + j * args.slot_x, tmx, &full_bar[s], c0 + j * 64, w0 + grp.s0,
                      n0, h0 + grp.r0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one chain per tap, B = shifted view of the halo tile =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      const int k_steps = args.brick_rows / kUmmaK;
      const int gpr = args.box_w / 8;                      // 8-pixel K groups per (h, n) row of the brick
      const int xw = args.box_w + grp.ns - 1;              // halo row width in pixels
      const uint32_t sbo_x = gpr == 1 ? xw * 128 : 1024;
      int ring_s = 0;
      uint32_t ring_phase = 0;
      for (int i = 0; i < num_bricks; ++i) {
        const int s = ring_s;
        const uint32_t phase = ring_phase;
        if (++ring_s == args.stages) { ring_s = 0; ring_phase ^= 1; }
        mbar_wait(&full_bar[s], phase, 32);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t x_addr = a_addr + kABlocks * args.slot_a;
        if (BN <= 128 && grp.ns == 1 && box_taps > 1 && grp.mask == (1 << box_taps) - 1) {
          // The taps of a filter column become consecutive 64-wide N blocks of ONE wide MMA: tap dr's view of the X
          // brick starts box_n * xw rows after tap dr-1's, which is exactly the descriptor's leading-dimension
          // stride.  N = taps * 64 (<= 256) amortises the A-operand fetch and the issue slot over all taps; a
          // 128-channel tile is two such MMAs (one per 64-channel block), accumulators laid out [block][tap][64].
          const uint32_t idesc_wide = make_idesc_bf16(kBlockM, 64 * box_taps, true, true) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);
          const uint32_t tap_pitch = args.box_n * xw * 128;
          // K step k covers the 8-pixel groups 2k, 2k+1 of the brick: (row, column group) advance with counters --
          // a runtime division per MMA on the issuing lane costs more than the MMA itself
          int row = 0, cg = 0;
          uint64_t adesc = make_smem_desc(a_addr, args.slot_a, 1024);
          uint64_t bdesc0[kBBlocks];
#pragma unroll
          for (int blk = 0; blk < kBBlocks; ++blk) bdesc0[blk] = make_smem_desc(x_addr + blk * args.slot_x, tap_pitch, sbo_x);
          for (int k = 0; k < k_steps; ++k) {
            const uint32_t x_off = (row * xw + cg * 8) * 128;
            cg += 2;
            while (cg >= gpr) { cg -= gpr; ++row; }
#pragma unroll
            for (int blk = 0; blk < kBBlocks; ++blk)
              mma_bf16_ss(tmem_base + blk * 64 * box_taps, adesc, desc_advance(bdesc0[blk], x_off), idesc_wide,
                          (i | k) != 0 ? 1u : 0u);
            adesc = desc_advance(adesc, kUmmaK * 128);
          }
        } else {
          int acc_idx = 0, dr = 0, ds = 0;
          for (int t = 0; t < box_taps; ++t) {
            const int dr_t = dr, ds_t = ds;                   // tap t = (dr, ds) of the group, row-major
            if (++ds == grp.ns) { ds = 0; ++dr; }
            if (!((grp.mask >> t) & 1)) continue;
            const uint32_t tmem_d = tmem_base + acc_idx * BN;
            ++acc_idx;
            int row = 0, cg = 0;                              // (h, n) row of the brick, 8-pixel column group
            for (int k = 0; k < k_steps; ++k) {
              const uint32_t x_off = ((row + dr_t * args.box_n) * xw + cg * 8 + ds_t) * 128;
              cg += 2;
              while (cg >= gpr) { cg -= gpr; ++row; }
              const uint64_t adesc = make_smem_desc(a_addr + k * (kUmmaK * 128), args.slot_a, 1024);
              const uint64_t bdesc = make_smem_desc(x_addr + x_off, args.slot_x, sbo_x);
              mma_bf16_ss(tmem_d, adesc, bdesc, kIdesc, (i | k) != 0 ? 1u : 0u);
            }
          }
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(accum_bar);
    }
  } else {
    // ===== epilogue: red.add every tap's 128 x BN accumulator into dW[m, tap, c] =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(accum_bar, 0, 33);
    tc_fence_after();
    float* dst_base = args.slab != nullptr ? args.slab + (long long)blockIdx.z * args.slab_elems : args.dw;
    int acc_idx = 0;
    for (int t = 0; t < box_taps; ++t) {
      if (!((grp.mask >> t) & 1)) continue;
      const int dr = t / grp.ns, ds = t - dr * grp.ns;
      const int tap = (grp.r0 + dr) * args.taps_w + grp.s0 + ds;
      float* out_row = dst_base + (static_cast<long long>(m0 + row) * args.num_taps + tap) * args.c_total + c0;
      const bool merged = BN <= 128 && grp.ns == 1 && box_taps > 1 && grp.mask == (1 << box_taps) - 1;
      const uint32_t tmem_t = tmem_base + acc_idx * BN;
      ++acc_idx;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        // merged layout: [64-channel block][tap][64]; per-tap layout: [tap][BN]
        const uint32_t col = merged ? (uint32_t)((c >> 1) * 64 * box_taps + t * 64 + (c & 1) * 32)
                                    : (uint32_t)((acc_idx - 1) * BN + c * 32);
        (void)tmem_t;
        tmem_ld_32x32(tmem_base + col + (static_cast<uint32_t>(quad * 32) << 16), v);
        tmem_ld_wait();
        float* o = out_row + c * 32;
        if (args.slab != nullptr || args.plain) {
          // every (split, tap, tile) slot of the slab is written exactly once: zeros when this split had no bricks
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(o + j) = num_bricks > 0 ? make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3])
                                                              : make_uint4(0u, 0u, 0u, 0u);
        } else if (num_bricks > 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// out[i] (+)= sum over splits of slab[s][i]; 128-bit accesses, 4 independent loads in flight per thread.
__global__ void __launch_bounds__(256)
reduce_slabs_kernel(const float4* __restrict__ slab, float4* __restrict__ out, long long nvec, int splits,
                    int accumulate) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long long)gridDim.x * 256) {
    float4 acc = accumulate ? out[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
      const float4 a = slab[(long long)s * nvec + i], b = slab[(long long)(s + 1) * nvec + i];
      const float4 c = slab[(long long)(s + 2) * nvec + i], d = slab[(long long)(s + 3) * nvec + i];
      acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; s < splits; ++s) {
      const float4 a = slab[(long long)s * nvec + i];
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    out[i] = acc;
  }
}

template <int BN>
void launch_impl(const WgradProblem& p, cudaStream_t stream) {
  MtArgs a{};
  a.fp16 = p.fp16;
  a.box_w = p.box_w; a.box_h = p.box_h; a.box_n = p.box_n;
  a.brick_rows = p.box_w * p.box_h * p.box_n;
  a.tiles_w = (p.out_w + p.box_w - 1) / p.box_w;
  a.tiles_h = (p.out_h + p.box_h - 1) / p.box_h;
  a.tiles_n = (p.out_n + p.box_n - 1) / p.box_n;
  a.taps_w = p.taps_w; a.num_taps = p.taps_h * p.taps_w;
  a.c_total = p.c_total; a.dw = p.dw;
  a.slab = p.slab;
  a.plain = (p.slab == nullptr && p.split_k <= 1 && !p.accumulate) ? 1 : 0;
  a.slab_elems = (long long)p.m_total * a.num_taps * p.c_total;
  const int gpr = p.box_w / 8;
  if (p.box_w % 8 != 0 || (gpr != 1 && gpr % 2 != 0) || a.brick_rows % kUmmaK != 0 || a.brick_rows > 128 ||
      p.m_total % kBlockM != 0 || p.c_total % BN != 0)
    throw std::runtime_error("wgrad_multitap: need box_w % 8 == 0 (1 or an even number of 8-pixel groups), "
                             "brick % 16 == 0, brick <= 128, M % 128 == 0, C % BLOCK_N == 0");
  // tap groups: as many taps per CTA as fit 512 TMEM columns.  When whole filter rows fit, taps are dealt out
  // evenly in row-major order (e.g. 9 taps, 8 fit -> 5 + 4) and a group's halo box spans the rows it touches;
  // otherwise a row is cut into column ranges.
  const int max_taps = 512 / BN;
  const int total_taps = p.taps_h * p.taps_w;
  int max_group_taps = 0, max_halo = 0;
  uint32_t shapes[2][2] = {{0, 0}, {0, 0}};
  int num_shapes = 0;
  auto add_group = [&](TapGroup g) {
    g.halo_rows = (p.box_h + g.nr - 1) * p.box_n * (p.box_w + g.ns - 1);
    g.map = -1;
    for (int k = 0; k < num_shapes; ++k)
      if (shapes[k][0] == (uint32_t)g.nr && shapes[k][1] == (uint32_t)g.ns) g.map = k;
    if (g.map < 0) {
      if (num_shapes == 2) throw std::runtime_error("wgrad_multitap: more than two halo shapes");
      shapes[num_shapes][0] = g.nr; shapes[num_shapes][1] = g.ns;
      g.map = num_shapes++;
    }
    if (a.num_groups == kMaxGroups) throw std::runtime_error("wgrad_multitap: too many tap groups");
    a.groups[a.num_groups++] = g;
    max_group_taps = std::max(max_group_taps, __builtin_popcount(g.mask));
    max_halo = std::max(max_halo, g.halo_rows);
  };
  // One group per filter COLUMN s: its taps (r = 0..kh-1, split when they do not fit TMEM) share an X brick with a
  // row halo only, so every tap view starts on a 1024-byte swizzle-atom boundary (ds = 0).  Pixel-shifted views
  // are legal but an 8-row group straddling two atoms doubles the operand fetches (see conv_gemm_halo.cu).
  (void)total_taps;
  const int rows_per_group = std::max(1, std::min(p.taps_h, max_taps));
  for (int s0 = 0; s0 < p.taps_w; ++s0)
    for (int r0 = 0; r0 < p.taps_h; r0 += rows_per_group) {
      TapGroup g{};
      g.r0 = r0; g.nr = std::min(rows_per_group, p.taps_h - r0);
      g.s0 = s0; g.ns = 1;
      g.mask = (1 << g.nr) - 1;
      add_group(g);
    }
  if (max_halo > 256) throw std::runtime_error("wgrad_multitap: halo brick exceeds a 256-row TMA box");
  a.slot_a = round_up(a.brick_rows * 128, 1024);
  a.slot_x = round_up(max_halo * 128, 1024);
  const int stage_bytes = (kBlockM / 64) * a.slot_a + (BN / 64) * a.slot_x;
  a.stages = std::min(8, (kMaxSmem - 2048) / stage_bytes);
  if (a.stages < 2) throw std::runtime_error("wgrad_multitap: stage does not fit shared memory twice");
  const size_t smem_bytes = 1024 + (size_t)a.stages * stage_bytes + (2 * a.stages + 1) * 8 + 16;

  // (C, W, N, H) dimension order => bricks land in shared memory as [h][n][w] rows of 64 channels
  auto make_map = [&](const void* base, int c, int w, int h, int n, int bw, int bh, int bn) {
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)n, (uint64_t)h};
    const uint64_t strides[3] = {(uint64_t)c * 2, (uint64_t)c * w * h * 2, (uint64_t)c * w * 2};
    const uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bn, (uint32_t)bh};
    return make_tensor_map_bf16(base, 4, dims, strides, box);
  };
  CUtensorMap tmap_dy = make_map(p.dy, p.m_total, p.out_w, p.out_h, p.out_n, p.box_w, p.box_h, p.box_n);
  CUtensorMap tmap_x[2];
  for (int k = 0; k < 2; ++k) {
    const int kk = k < num_shapes ? k : 0;
    tmap_x[k] = make_map(p.x, p.c_total, p.in_w, p.in_h, p.out_n, p.box_w + (int)shapes[kk][1] - 1,
                         p.box_h + (int)shapes[kk][0] - 1, p.box_n);
  }
  dim3 grid(p.m_total / kBlockM, a.num_groups * (p.c_total / BN), p.split_k > 1 ? p.split_k : 1);
  const int cols = max_group_taps * BN;
  auto run = [&](auto kernel) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    launch_pdl(kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, tmap_dy, tmap_x[0], tmap_x[1], a);
  };
  if (cols <= 64) run(wgrad_multitap_kernel<BN, 64>);
  else if (cols <= 128) run(wgrad_multitap_kernel<BN, 128>);
  else if (cols <= 256) run(wgrad_multitap_kernel<BN, 256>);
  else run(wgrad_multitap_kernel<BN, 512>);
  DDPT_CUDA_CHECK(cudaGetLastError());
  if (p.slab != nullptr) {
    const long long nvec = a.slab_elems / 4;
    long long want = (nvec + 255) / 256;
    const int blocks = (int)std::min<long long>(want, (long long)sm_count() * 8);
    launch_pdl(reduce_slabs_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const float4*>(p.slab),
                                                    reinterpret_cast<float4*>(p.dw), nvec, (int)grid.z, p.accumulate);
    DDPT_CUDA_CHECK(cudaGetLastError());
  }
}

}  // namespace

// Same contract as launch_wgrad_gemm, except that the output is always accumulated (red.add): the caller
// provides a zero-initialised (or beta = 1) dW.
void launch_wgrad_gemm_multitap(const WgradProblem& p, int block_n, cudaStream_t stream) {
  switch (block_n) {
    case 64: return launch_impl<64>(p, stream);
    case 128: return launch_impl<128>(p, stream);
    case 256: return launch_impl<256>(p, stream);
  }
  throw std::runtime_error("wgrad_multitap: block_n must be 64, 128 or 256");
}

}  // namespace ddpt
