// Halo-reuse implicit-GEMM convolution (forward and data-gradient) for sm_100a: tcgen05 + TMEM + TMA.
//
// conv_gemm_persistent.cu streams one activation brick per (filter tap, 64-channel chunk): every input pixel
// crosses L2->SM nine times and the layer is L2-bandwidth bound when C_in / C_out are small (Toy_Net conv.3).
// Here a brick is loaded once per (channel chunk, filter COLUMN) together with its row halo,
//     [(bh + kh - 1) rows][bn images][8 pixels][64 channels]                (TMA box over a (C, W, N, H) map)
// and the kh taps of that filter column are *descriptor views* into the tile: the start address moves by whole
// 1024-byte swizzle atoms (one 8-pixel group per (row, image)), so every view stays atom-aligned.  (Views shifted
// by single pixels also compute correctly -- tcgen05.mma swizzles on absolute shared-memory addresses,
// profiles/umma_descriptor_probe.txt -- but an 8-row group that straddles two atoms costs two shared-memory
// fetches and the kernel becomes operand-fetch bound; measured in profiles/.)  Activation traffic drops from
// kh*kw to kw*(bh+kh-1)/bh bricks per chunk (3.75 instead of 9 for 3x3, bh = 8).  Weights either stay resident in
// shared memory for the whole kernel (when all taps x chunks x BLOCK_N fit: conv.3) or stream through their
// own ring.  Output tiles are 16 groups of 8 consecutive pixels ordered (h, n, w); everything else -- warp
// roles, double-buffered TMEM accumulator, bias / ReLU / BatchNorm-statistics epilogue -- matches the
// persistent kernel.
#include <cuda.h>
#include <cstdlib>
#include <stdexcept>
#include "../include/ddpt_nn.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"
#include "tensor_map.h"

namespace ddpt {
namespace {

using namespace umma;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
// Three warpgroups: warps 0..3 control (0: TMA producer, 1: MMA issuer, 2: second issuer when NI == 2, 3: idle), warps
// 4..11 epilogue.  After the setup the control group gives registers back (setmaxnreg.dec) and the epilogue groups take
// them (setmaxnreg.inc): the epilogue's per-thread BatchNorm accumulators need ~200 registers, the single-lane roles ~50.
constexpr int kThreads = 384;
constexpr int kEpilogueWarps = 8;
constexpr int kFirstEpilogueWarp = 4;
constexpr int kControlRegs = 104, kEpilogueRegs = 200;     // 128 x 104 + 256 x 200 = 64512 <= 65536
constexpr int kMaxSmem = 227 * 1024;

struct HaloArgs {
  int tiles_w, tiles_h;
  int box_h, box_n;                 // box_w is 8
  int out_w, out_h, out_n;
  int taps_w, taps_h, tap_sign;
  int org_h_off, org_w_off;         // halo tile origin relative to the tile's (h0, w0)
  int k_chunks, n_total;
  int a_slot_bytes, a_tile_bytes;
  int a_stages, b_stages;
  long long ld_out;
  void* out;
  const float* bias;
  float* stat_partials;
  double* stat_accum;
  int relu, out_fp32;
  int fp16;                         // operands and 16-bit outputs are fp16 instead of bf16
};

// Column sums over the 32 rows a warp holds (one row per lane, 32 columns per lane): a transposing butterfly -- after
// log2(32) exchange levels lane L holds, in element 0, the sum of column L over all 32 lanes (31 shuffles per array).
__device__ __forceinline__ void warp_column_sums(float (&s1)[32], float (&s2)[32], int lane) {
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
}

// NI = 2 (resident weights, BN <= 128): TWO MMA-issuing warps.  A 128 x 64 x 16 MMA is 32 tensor cycles, less than one
// in-order thread needs to build and issue it (profiles/ncu_gemm_v2.txt: conv.3 data gradient, tensor pipe 39 % active,
// the issuing lane 65 % busy with descriptor arithmetic + UTCHMMA).  Issuer j takes every second activation brick
// (the 3 filter-row taps x 4 K-steps of one halo tile) and accumulates into its OWN TMEM accumulator; the epilogue adds
// the two.  Both issuers run on different SM sub-partitions, so descriptor math proceeds in parallel while the tensor
// pipe executes their MMAs in arrival order.
//
// KC > 0 (resident weights, 3x3 filter, KC channel chunks): the issuer's loops over (chunk, filter column, filter row,
// K step) are fully unrolled with compile-time trip counts.  Every weight-tile descriptor is then `base + immediate`
// and an activation view `brick base + dr * stride + k * 32 B`: two integer adds per MMA instead of the ~15-20
// instructions of runtime tap arithmetic and descriptor encoding the generic loop spends -- on 64-column tiles (32
// tensor cycles per MMA) that arithmetic, not the tensor pipe, set the pace (tensor pipe 39 % active).
template <int BN, bool B_MN, bool B_RES, int NI, int KC>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const HaloArgs args, const int num_m_tiles, const int num_n_tiles) {
  constexpr int kBStageBytes = BN * kBlockK * 2;
  const uint32_t kIdesc = make_idesc_bf16(kBlockM, BN, false, B_MN) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);
  constexpr uint32_t kTmemCols = 2 * BN * NI;
  static_assert(NI == 1 || (NI == 2 && B_RES && BN <= 128), "two issuers: resident weights, 2 x 2 x BN TMEM columns");
  static_assert(KC == 0 || (B_RES && NI == 1), "unrolled issuer: resident weights, single issuer");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_taps = args.taps_h * args.taps_w;
  const int total_kb = num_taps * args.k_chunks;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int b_slots = B_RES ? total_kb : args.b_stages;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + args.a_stages * args.a_slot_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + b_slots * kBStageBytes);
  uint64_t* a_empty = a_full + args.a_stages;
  uint64_t* b_full = a_empty + args.a_stages;
  uint64_t* b_empty = b_full + args.b_stages;
  uint64_t* tmem_full = b_empty + args.b_stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_ready = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < args.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < args.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], NI); mbar_init(&tmem_empty[s], kEpilogueWarps); }
    mbar_init(b_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // operands come from the previous kernel of the stream: wait for it only now, after the setup
  pdl_trigger();

  auto load_b = [&](uint8_t* dst, uint64_t* bar, int tap, int chunk, int col0) {
    if (!B_MN) {
      tma_load_2d(dst, &tmap_b, bar, (tap * args.k_chunks + chunk) * kBlockK, col0);
    } else {
#pragma unroll
      for (int j = 0; j < BN / 64; ++j)
        tma_load_2d(dst + j * (kBlockK * 128), &tmap_b, bar, tap * args.n_total + col0 + j * 64, chunk * kBlockK);
    }
  };

  if (warp < kFirstEpilogueWarp) {
  reg_dealloc<kControlRegs>();      // the whole control warpgroup, before its warps split into roles
  if (warp == 0) {
    // ===== TMA producer: one halo brick per (tile, channel chunk); weights resident or streamed per tap =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      if (B_RES) {
        mbar_arrive_expect_tx(b_ready, static_cast<uint32_t>(total_kb) * kBStageBytes);
        for (int tap = 0; tap < num_taps; ++tap)
          for (int chunk = 0; chunk < args.k_chunks; ++chunk)
            load_b(smem_b + (tap * args.k_chunks + chunk) * kBStageBytes, b_ready, tap, chunk, 0);
      }
      int a_s = 0, b_s = 0;
      uint32_t a_phase = 0, b_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tile_m = tile / num_n_tiles, tile_n = tile - tile_m * num_n_tiles;
        const int tw = tile_m % args.tiles_w;
        const int th = (tile_m / args.tiles_w) % args.tiles_h;
        const int tn = tile_m / (args.tiles_w * args.tiles_h);
        const int w0 = tw * 8, h0 = th * args.box_h, n0 = tn * args.box_n;
        for (int chunk = 0; chunk < args.k_chunks; ++chunk) {
          for (int cs = 0; cs < args.taps_w; ++cs) {       // one activation brick per filter column
            const int sa = a_s;
            const uint32_t a_ph = a_phase;
            if (++a_s == args.a_stages) { a_s = 0; a_phase ^= 1; }
            mbar_wait(&a_empty[sa], a_ph ^ 1, 41);
            mbar_arrive_expect_tx(&a_full[sa], args.a_tile_bytes);
            tma_load_4d(smem_a + sa * args.a_slot_bytes, &tmap_a, &a_full[sa], chunk * kBlockK,
                        w0 + args.org_w_off + cs, n0, h0 + args.org_h_off);
            if (!B_RES) {
              for (int dr = 0; dr < args.taps_h; ++dr) {
                const int r = args.tap_sign > 0 ? dr : args.taps_h - 1 - dr;
                const int sx = args.tap_sign > 0 ? cs : args.taps_w - 1 - cs;
                const int sb = b_s;
                const uint32_t b_ph = b_phase;
                if (++b_s == args.b_stages) { b_s = 0; b_phase ^= 1; }
                mbar_wait(&b_empty[sb], b_ph ^ 1, 42);
                mbar_arrive_expect_tx(&b_full[sb], kBStageBytes);
                load_b(smem_b + sb * kBStageBytes, &b_full[sb], r * args.taps_w + sx, chunk, tile_n * BN);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || (NI == 2 && warp == 2)) {
    // ===== MMA issuer(s): every tap is a shifted view of the halo brick =====
    const int me = warp == 1 ? 0 : 1;                       // issuer index (NI == 2: warp 2 is the second issuer)
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      int a_s = 0, b_s = 0;
      uint32_t a_phase = 0, b_phase = 0, t = 0, brick = 0;
      if (B_RES) mbar_wait(b_ready, 0, 45);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 44);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (acc * NI + me) * BN;
        if constexpr (KC > 0) {
          // fixed 3x3 x KC shape: everything but the ring slot is a compile-time constant
          constexpr int kSign = B_MN ? -1 : 1;              // data gradient (MN-major weights) walks the taps backwards
          const uint64_t bdesc_base = B_MN ? make_smem_desc(smem_u32(smem_b), kBlockK * 128, 1024)
                                           : make_smem_desc(smem_u32(smem_b), 0, 1024);
          const uint32_t dr_stride = static_cast<uint32_t>(args.box_n) * 1024u;
#pragma unroll
          for (int chunk = 0; chunk < KC; ++chunk) {
#pragma unroll
            for (int cs = 0; cs < 3; ++cs, ++brick) {
              const int sa = a_s;
              const uint32_t a_ph = a_phase;
              if (++a_s == args.a_stages) { a_s = 0; a_phase ^= 1; }
              mbar_wait(&a_full[sa], a_ph, 43);
              tc_fence_after();
              const uint64_t adesc_base = make_smem_desc(smem_u32(smem_a + sa * args.a_slot_bytes), 0, 1024);
#pragma unroll
              for (int dr = 0; dr < 3; ++dr) {
                constexpr int kDummy = 0; (void)kDummy;
                const int r = kSign > 0 ? dr : 2 - dr;
                const int sx = kSign > 0 ? cs : 2 - cs;
                const uint32_t b_off = static_cast<uint32_t>(((r * 3 + sx) * KC + chunk) * kBStageBytes);
                const uint64_t adesc_dr = desc_advance(adesc_base, dr * dr_stride);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  mma_bf16_ss(tmem_d, desc_advance(adesc_dr, k * (kUmmaK * 2)),
                              desc_advance(bdesc_base, b_off + k * (B_MN ? kUmmaK * 128 : kUmmaK * 2)), kIdesc,
                              (chunk | cs | dr | k) != 0 ? 1u : 0u);
              }
              mma_commit(&a_empty[sa]);
            }
          }
          mma_commit(&tmem_full[acc]);
          continue;
        }
        uint32_t started = 0;                               // this issuer's accumulator holds data of this tile
        for (int chunk = 0; chunk < args.k_chunks; ++chunk) {
          for (int cs = 0; cs < args.taps_w; ++cs, ++brick) {
            const int sa = a_s;
            const uint32_t a_ph = a_phase;
            if (++a_s == args.a_stages) { a_s = 0; a_phase ^= 1; }
            if (NI > 1 && (brick & 1u) != (uint32_t)me) continue;      // the other issuer's brick
            mbar_wait(&a_full[sa], a_ph, 43);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + sa * args.a_slot_bytes);
            for (int dr = 0; dr < args.taps_h; ++dr) {
              const int r = args.tap_sign > 0 ? dr : args.taps_h - 1 - dr;
              const int sx = args.tap_sign > 0 ? cs : args.taps_w - 1 - cs;
              const int tap = r * args.taps_w + sx;
              uint32_t b_addr;
              int sb = 0;
              if (B_RES) {
                b_addr = smem_u32(smem_b + (tap * args.k_chunks + chunk) * kBStageBytes);
              } else {
                sb = b_s;
                const uint32_t b_ph = b_phase;
                if (++b_s == args.b_stages) { b_s = 0; b_phase ^= 1; }
                mbar_wait(&b_full[sb], b_ph, 46);
                tc_fence_after();
                b_addr = smem_u32(smem_b + sb * kBStageBytes);
              }
              const uint32_t a_view = a_addr + (dr * args.box_n) * 1024;     // whole swizzle atoms: aligned view
              const uint64_t adesc0 = make_smem_desc(a_view, 0, 1024);
              const uint64_t bdesc0 = B_MN ? make_smem_desc(b_addr, kBlockK * 128, 1024) : make_smem_desc(b_addr, 0, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                mma_bf16_ss(tmem_d, desc_advance(adesc0, k * (kUmmaK * 2)),
                            desc_advance(bdesc0, k * (B_MN ? kUmmaK * 128 : kUmmaK * 2)), kIdesc,
                            (started | dr | k) != 0 ? 1u : 0u);
              if (!B_RES) mma_commit(&b_empty[sb]);
            }
            started = 1;
            mma_commit(&a_empty[sa]);
          }
        }
        mma_commit(&tmem_full[acc]);
      }
    }
  }
  } else {
    reg_alloc<kEpilogueRegs>();
    // ===== epilogue: row m = group (h, n) x 8 pixels; quadrant = warp & 3, column half = (warp - 4) >> 2 =====
    const int quad = warp & 3;
    const int half = (warp - kFirstEpilogueWarp) >> 2;
    const int row = quad * 32 + lane;
    const int grp = row >> 3, dw = row & 7;
    const int dh = grp / args.box_n, dn = grp - dh * args.box_n;
    constexpr int kChunksPerWarp = (BN / 32) / 2;
    float acc_s1[kChunksPerWarp], acc_s2[kChunksPerWarp];
#pragma unroll
    for (int cc = 0; cc < kChunksPerWarp; ++cc) acc_s1[cc] = acc_s2[cc] = 0.f;
    // BatchNorm statistics of forward tiles up to 128 columns wide: every thread keeps running sums of ITS row's values
    // for its 32 / 64 columns across all tiles of the CTA, and the cross-lane fold runs once per CTA at the end.  The
    // per-tile butterfly (124 shuffles + 250 selects per warp and tile) made conv.3's forward EPILOGUE-bound: 42 M warp
    // instructions, shuffle pipe saturated, 7200 cycles per tile against 2304 cycles of MMAs, tensor pipe 33 % busy
    // (profiles/ncu_r2_fwd3_halo_step.txt).  128 fp32 accumulators fit next to the 32 live tile values (200 registers after setmaxnreg).
    constexpr bool kRegStats = !B_MN && BN <= 128 && NI == 1;
    constexpr int kRegCols = kRegStats ? kChunksPerWarp * 32 : 1;
    float racc1[kRegCols], racc2[kRegCols];
#pragma unroll
    for (int j = 0; j < kRegCols; ++j) racc1[j] = racc2[j] = 0.f;
    const bool want_stats = args.stat_partials != nullptr || args.stat_accum != nullptr;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int tile_m = tile / num_n_tiles, tile_n = tile - tile_m * num_n_tiles;
      const int tw = tile_m % args.tiles_w;
      const int th = (tile_m / args.tiles_w) % args.tiles_h;
      const int tn = tile_m / (args.tiles_w * args.tiles_h);
      const int n = tn * args.box_n + dn, h = th * args.box_h + dh, w = tw * 8 + dw;
      const bool valid = dh < args.box_h && n < args.out_n && h < args.out_h && w < args.out_w;
      const long long out_row = (static_cast<long long>(n) * args.out_h + h) * args.out_w + w;
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase, 47);
      tc_fence_after();
#pragma unroll (kRegStats ? kChunksPerWarp : 1)
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int c = half * kChunksPerWarp + cc;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (acc * NI) * BN + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (NI > 1) {                                          // second issuer's partial sums
          tmem_ld_32x32(tmem_base + (acc * NI + 1) * BN + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
        }
        const int col = tile_n * BN + c * 32;
        if (args.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(args.bias + col + j);
        }
        if (want_stats) {
          if constexpr (kRegStats) {
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                racc1[cc * 32 + j] += f[j];
                racc2[cc * 32 + j] = fmaf(f[j], f[j], racc2[cc * 32 + j]);
              }
            }
          } else {
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s1[j] = valid ? f[j] : 0.f;
              s2[j] = s1[j] * s1[j];
            }
            warp_column_sums(s1, s2, lane);
            acc_s1[cc] += s1[0];
            acc_s2[cc] += s2[0];
          }
        }
        if (args.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (valid) {
          if (args.out_fp32) {
            float* o = static_cast<float*>(args.out) + out_row * args.ld_out + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if constexpr (kRegStats) {
      if (want_stats) {
#pragma unroll
        for (int cc = 0; cc < kChunksPerWarp; ++cc) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1[j] = racc1[cc * 32 + j]; s2[j] = racc2[cc * 32 + j]; }
          warp_column_sums(s1, s2, lane);
          acc_s1[cc] = s1[0];
          acc_s2[cc] = s2[0];
        }
      }
    }
    if (args.stat_accum != nullptr) {
      const int my_tile_n = blockIdx.x % num_n_tiles;
#pragma unroll
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int col = my_tile_n * BN + (half * kChunksPerWarp + cc) * 32 + lane;
        atomicAdd(args.stat_accum + col, (double)acc_s1[cc]);
        atomicAdd(args.stat_accum + args.n_total + col, (double)acc_s2[cc]);
      }
    } else if (args.stat_partials != nullptr) {
      const int my_tile_n = blockIdx.x % num_n_tiles;
      float* dst = args.stat_partials + (static_cast<long long>(blockIdx.x) * 4 + quad) * 2 * args.n_total;
      for (int tn2 = 0; tn2 < num_n_tiles; ++tn2) {
#pragma unroll
        for (int cc = 0; cc < kChunksPerWarp; ++cc) {
          const int col = tn2 * BN + (half * kChunksPerWarp + cc) * 32 + lane;
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

int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <int BN, bool B_MN>
int launch_impl(const ConvGemmProblem& p, cudaStream_t stream) {
  constexpr int kBStageBytes = BN * kBlockK * 2;
  if (p.box_w != 8 || p.a_c % kBlockK != 0 || p.n_total % BN != 0 || p.box_h * p.box_n > 16 || p.split_k > 1)
    throw std::runtime_error("conv_gemm_halo: needs box_w == 8, box_h*box_n <= 16, C % 64 == 0, N % BLOCK_N == 0, no split-K");
  HaloArgs a{};
  a.box_h = p.box_h; a.box_n = p.box_n;
  a.out_w = p.out_w; a.out_h = p.out_h; a.out_n = p.out_n;
  a.tiles_w = (p.out_w + 7) / 8;
  a.tiles_h = (p.out_h + p.box_h - 1) / p.box_h;
  const int tiles_n = (p.out_n + p.box_n - 1) / p.box_n;
  a.taps_w = p.taps_w; a.taps_h = p.taps_h; a.tap_sign = p.tap_sign;
  a.org_h_off = p.a_h_off - (p.tap_sign > 0 ? 0 : p.taps_h - 1);
  a.org_w_off = p.a_w_off - (p.tap_sign > 0 ? 0 : p.taps_w - 1);
  a.k_chunks = p.a_c / kBlockK;
  a.n_total = p.n_total;
  const int halo_rows = (p.box_h + p.taps_h - 1) * p.box_n * 8;
  a.a_tile_bytes = halo_rows * 128;
  // the 16-group view may read past the loaded rows when box_h * box_n < 16: keep that inside the slot
  const int view_rows = (16 + (p.taps_h - 1) * p.box_n) * 8;
  a.a_slot_bytes = round_up((view_rows > halo_rows ? view_rows : halo_rows) * 128, 1024);
  a.ld_out = p.ld_out; a.out = p.out; a.bias = p.bias; a.stat_partials = p.stat_partials; a.stat_accum = p.stat_accum;
  a.relu = p.relu; a.out_fp32 = p.out_fp32; a.fp16 = p.fp16;

  const int total_kb = p.taps_h * p.taps_w * a.k_chunks;
  const long long b_res_bytes = (long long)total_kb * kBStageBytes;
  int budget = kMaxSmem - 2048;
  const bool resident = p.n_total == BN && b_res_bytes + 3 * a.a_slot_bytes <= budget && b_res_bytes < (1 << 20);
  if (resident) {
    a.b_stages = 1;                       // unused ring, one dummy barrier pair
    a.a_stages = std::min(6, (int)((budget - b_res_bytes) / a.a_slot_bytes));
  } else {
    // split the budget roughly 1:1 between activation bricks and weight tiles, at least 3 bricks in flight
    a.a_stages = std::max(3, std::min(5, (budget / 2) / a.a_slot_bytes));
    a.b_stages = std::min(8, (budget - a.a_stages * a.a_slot_bytes) / kBStageBytes);
  }
  if (a.a_stages < 2 || a.b_stages < 1) throw std::runtime_error("conv_gemm_halo: tile does not fit shared memory");
  const size_t smem_bytes = 1024 + (size_t)a.a_stages * a.a_slot_bytes +
                            (resident ? (size_t)b_res_bytes : (size_t)a.b_stages * kBStageBytes) +
                            (2 * a.a_stages + 2 * a.b_stages + 5) * 8 + 16;
  if (smem_bytes > (size_t)kMaxSmem) throw std::runtime_error("conv_gemm_halo: shared memory budget exceeded");

  // A: (C, W, N, H) map, box (64, 8, box_n, box_h + taps_h - 1): row-halo brick, rows ordered [h][n][w]
  const uint64_t a_dims[4] = {(uint64_t)p.a_c, (uint64_t)p.a_w, (uint64_t)p.a_n, (uint64_t)p.a_h};
  const uint64_t a_strides[3] = {(uint64_t)p.a_c * 2, (uint64_t)p.a_c * p.a_w * p.a_h * 2, (uint64_t)p.a_c * p.a_w * 2};
  const uint32_t a_box[4] = {(uint32_t)kBlockK, 8u, (uint32_t)p.box_n, (uint32_t)(p.box_h + p.taps_h - 1)};
  CUtensorMap tmap_a = make_tensor_map_bf16(p.a, 4, a_dims, a_strides, a_box);
  CUtensorMap tmap_b;
  const int num_taps = p.taps_h * p.taps_w;
  if (!B_MN) {
    const uint64_t b_dims[2] = {(uint64_t)num_taps * p.a_c, (uint64_t)p.n_total};
    const uint64_t b_strides[1] = {(uint64_t)p.b_ld * 2};
    const uint32_t b_box[2] = {(uint32_t)kBlockK, (uint32_t)BN};
    tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);
  } else {
    const uint64_t b_dims[2] = {(uint64_t)num_taps * p.n_total, (uint64_t)p.a_c};
    const uint64_t b_strides[1] = {(uint64_t)p.b_ld * 2};
    const uint32_t b_box[2] = {64u, (uint32_t)kBlockK};
    tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);
  }
  const int tiles_m = a.tiles_w * a.tiles_h * tiles_n;
  const int tiles_n_cols = p.n_total / BN;
  const int total_tiles = tiles_m * tiles_n_cols;
  int grid = total_tiles < sm_count() ? total_tiles : sm_count();
  grid -= grid % tiles_n_cols;
  if (grid < tiles_n_cols) grid = tiles_n_cols;
  auto run = [&](auto kernel, int threads) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    launch_pdl(kernel, dim3(grid), dim3(threads), smem_bytes, stream, tmap_a, tmap_b, a, tiles_m, tiles_n_cols);
  };
  // Two issuing warps (DDPT_HALO_ISSUERS=2, opt-in): resident weights, at least two activation bricks per tile.  Measured
  // on conv.3's data gradient (64 columns): 192 us with two issuers on an even 4-stage ring vs 167 us with one issuer on
  // a 5-stage ring -- the in-order tensor pipe gains nothing from a second instruction stream once the first one's
  // descriptor math is a single add (desc_advance), and the ring loses a stage.  Kept for experiments.
  static const bool dual_ok = std::getenv("DDPT_HALO_ISSUERS") != nullptr && std::atoi(std::getenv("DDPT_HALO_ISSUERS")) == 2;
  if constexpr (BN <= 128) {
    if (resident && dual_ok && a.k_chunks * p.taps_w >= 2 && a.a_stages >= 2) {
      // an EVEN ring: stage s then always belongs to issuer s % 2, which therefore observes every phase of that stage's
      // barriers.  With an odd ring an issuer would see only every second phase of a stage, and a parity wait cannot tell
      // "my brick has landed" from "the brick two laps ago had" when it runs ahead of the other issuer (seen as a sporadic
      // launch failure in the first version).
      a.a_stages &= ~1;
      run(conv_gemm_halo_kernel<BN, B_MN, true, 2, 0>, kThreads);
      DDPT_CUDA_CHECK(cudaGetLastError());
      return p.stat_partials != nullptr ? grid : tiles_m;
    }
  }
  // fixed-shape issuer (3x3 filter, tap direction matching the operand layout, 1 or 2 channel chunks): DDPT_HALO_UNROLL=0
  // keeps the generic loops
  static const bool unroll_ok = std::getenv("DDPT_HALO_UNROLL") == nullptr || std::atoi(std::getenv("DDPT_HALO_UNROLL")) != 0;
  const bool fixed = resident && unroll_ok && p.taps_h == 3 && p.taps_w == 3 && p.tap_sign == (B_MN ? -1 : 1);
  if constexpr (BN <= 128) {
    if (fixed && a.k_chunks == 1) { run(conv_gemm_halo_kernel<BN, B_MN, true, 1, 1>, kThreads); DDPT_CUDA_CHECK(cudaGetLastError()); return p.stat_partials != nullptr ? grid : tiles_m; }
    if (fixed && a.k_chunks == 2) { run(conv_gemm_halo_kernel<BN, B_MN, true, 1, 2>, kThreads); DDPT_CUDA_CHECK(cudaGetLastError()); return p.stat_partials != nullptr ? grid : tiles_m; }
  }
  if (resident) run(conv_gemm_halo_kernel<BN, B_MN, true, 1, 0>, kThreads);
  else run(conv_gemm_halo_kernel<BN, B_MN, false, 1, 0>, kThreads);
  DDPT_CUDA_CHECK(cudaGetLastError());
  return p.stat_partials != nullptr ? grid : tiles_m;
}

}  // namespace

int launch_conv_gemm_halo(const ConvGemmProblem& p, int block_n, cudaStream_t stream) {
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
  throw std::runtime_error("conv_gemm_halo: block_n must be 64, 128 or 256");
}

}  // namespace ddpt
