// Persistent implicit-GEMM convolution / linear forward and data-gradient for sm_100a.
//
// Replaces cuDNN convolution fwd/dgrad and cuBLAS linear fwd/dgrad of the reference stack
// (Toy_Net layers K3/K6/K8/K11/K13).  One CTA per SM walks a static list of
// 128 x BLOCK_N output tiles; the accumulator is double-buffered in TMEM so the epilogue of tile i
// overlaps the TMA/MMA main loop of tile i+1 and the one-time setup (barriers, TMEM allocation,
// descriptor prefetch) is paid once per SM instead of once per tile:
//
//   warp 0      TMA producer: per K block (one filter tap x 64 input channels) one 4-D box load of
//               the activation brick shifted by the tap, plus the matching weight tile
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per K block),
//               tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, each owning half of the columns): tcgen05.ld 32 columns at a time, + bias, ReLU, BatchNorm partial
//               statistics (column sum / sum of squares per tile), bf16 or fp32 store
//
// Operands are staged in 128-byte-swizzled shared memory exactly as TMA writes them; the matrix
// descriptors describe that layout (K-major for activations, K- or MN-major for weights).
#include <cuda.h>
#include "../include/ddpt_nn.h"
#include "../include/ddpt_common.cuh"
#include "../include/umma.cuh"
#include "tensor_map.h"
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace ddpt {
namespace {

using namespace umma;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 320;               // producer, mma, 8 epilogue warps
constexpr int kEpilogueWarps = 8;
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
  int stages;                   // pipeline depth (stage = kbs K blocks of A [+ B unless resident])
  int kbs;                      // K blocks (tap x 64-channel chunk) per pipeline stage: amortises the barrier round trip
  int a_w, a_h;                 // extent of the activation tensor (taps whose whole brick is out of range are skipped)
  int pair_interleave;          // CTA pair: the two row tiles of a unit are the same pixel brick of adjacent image groups
  const int* schedule;          // optional host-planned work lists: schedule[cta * sched_len + k] = k-th unit of CTA (pair)
  int sched_len;                // `cta`, -1 terminated; null: round robin (unit = cta + k * #ctas)
  int fp16;                     // operands and 16-bit outputs are fp16 instead of bf16
};

// k-th work unit of this CTA (pair), or -1 when it has none left.  Data-gradient tiles cost between 1 and taps_h x taps_w
// filter taps once out-of-range taps are skipped, so a round-robin walk leaves some CTAs with 20-30 % more work than
// the average; the launcher plans the lists with a longest-processing-time-first heuristic instead.
__device__ __forceinline__ int work_unit(const KernelArgs& a, int cta, int num_ctas, int num_units, int k) {
  if (a.schedule != nullptr) return k < a.sched_len ? __ldg(a.schedule + cta * a.sched_len + k) : -1;
  const int unit = cta + k * num_ctas;
  return unit < num_units ? unit : -1;
}


// Row tile -> brick coordinates.  Default order: brick column fastest, then brick row, then image group.  With
// pair_interleave the two tiles of a CTA pair (tile_m = 2q, 2q+1) share the pixel position and take adjacent image
// groups, so both need exactly the same filter taps (data gradient: border positions skip out-of-range taps).
__host__ __device__ __forceinline__ void decode_tile(const KernelArgs& a, int tile_m, int& tw, int& th, int& tn) {
  const int positions = a.tiles_w * a.tiles_h;
  int pos;
  if (a.pair_interleave) {
    const int q = tile_m >> 1;
    pos = q % positions;
    tn = (q / positions) * 2 + (tile_m & 1);
  } else {
    pos = tile_m % positions;
    tn = tile_m / positions;
  }
  tw = pos % a.tiles_w;
  th = pos / a.tiles_w;
}

// Filter taps whose shifted brick intersects the activation tensor at all: [lo, hi] along one axis.  For the forward
// ('valid') convolution every tap qualifies; for the data gradient a brick at the border of the input image only
// sees part of the filter -- with one-pixel bricks (128 images x 1 x 1) the MMA count drops from taps x pixels to the
// nominal sum over pixels of the taps that actually reach them (Toy_Net conv.10: 576 instead of 900 per image group).
__host__ __device__ __forceinline__ void tap_range(int origin, int box, int extent, int taps, int sign, int& lo, int& hi) {
  lo = taps; hi = -1;
  for (int t = 0; t < taps; ++t) {
    const int start = origin + sign * t;
    if (start + box > 0 && start < extent) { lo = lo < t ? lo : t; hi = hi > t ? hi : t; }
  }
  if (hi < 0) { lo = 0; hi = 0; }            // nothing in range (padding tile): one tap of zeros defines the accumulator
}

struct TileTaps { int tx0, tx1, ty0, ty1; };
// Union over the MC row tiles of a work unit (the leader issues one MMA stream for both CTAs).
template <int MC>
__host__ __device__ __forceinline__ TileTaps unit_taps(const KernelArgs& a, int tile_m0) {
  TileTaps r{a.taps_w, -1, a.num_taps / a.taps_w, -1};
#pragma unroll
  for (int i = 0; i < MC; ++i) {
    int tw, th, tn, lo, hi;
    decode_tile(a, tile_m0 + i, tw, th, tn);
    tap_range(tw * a.box_w + a.a_w_off, a.box_w, a.a_w, a.taps_w, a.tap_sign, lo, hi);
    r.tx0 = r.tx0 < lo ? r.tx0 : lo; r.tx1 = r.tx1 > hi ? r.tx1 : hi;
    tap_range(th * a.box_h + a.a_h_off, a.box_h, a.a_h, a.num_taps / a.taps_w, a.tap_sign, lo, hi);
    r.ty0 = r.ty0 < lo ? r.ty0 : lo; r.ty1 = r.ty1 > hi ? r.ty1 : hi;
  }
  return r;
}


// B_RES: the whole weight matrix (all taps x channel chunks of the single column block) is loaded into shared
// memory once per CTA and stays resident; only activation bricks stream through the stage ring.  Halves (fwd) /
// cuts by a third (dgrad) the L2->SM traffic of layers whose weights fit (Toy_Net conv.3: 144 KiB).
//
// MC = 2 (CTA pair, tcgen05 cta_group::2): two CTAs of a cluster -- two SMs of one TPC -- compute a 256 x BN tile
// with ONE paired MMA stream issued by the leader.  Each CTA stages its own 128 activation rows and HALF of the
// weight tile, so per K block a CTA's shared memory takes 32 KiB of TMA writes and 8 KiB of operand reads per MMA
// instead of 48 / 12 KiB at BN = 256.  Why: the single-CTA kernel plateaued at 70 % tensor-pipe activity with TMA
// writes (94 B/clk) + operand reads (96 B/clk) ~ 1.4x the 128 B/clk shared-memory port; TMA multicast of the weight
// tile alone (same bytes written, fewer L2 reads) measured +-0 % (profiles/README.md), i.e. the port, not L2, binds.
//   * both CTAs' TMA loads complete on the LEADER's full barrier (cta_group::2 loads, peer bit cleared)
//   * tcgen05.commit.cta_group::2 multicasts stage release / accumulator-ready onto both CTAs' barriers
//   * both CTAs' epilogue warps arrive on the leader's tmem_empty barrier (remote mbarrier arrive)
template <int BN, bool B_MN, bool B_RES, int MC>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                            const KernelArgs args, const int num_m_tiles, const int num_n_tiles) {
  constexpr int kBStageBytes = (BN / MC) * kBlockK * 2;     // one K block of B (this CTA's share of the columns)
  const uint32_t kIdesc = make_idesc_bf16(kBlockM * MC, BN, false, B_MN) ^ (args.fp16 ? kIdescBf16ToF16 : 0u);
  constexpr uint32_t kTmemCols = 2 * BN;                   // two accumulator stages

  // CTA pair: streamed weights; an MN-major weight tile is loaded in 64-column boxes, so a pair needs BN >= 128 there.
  // K-major weights split at any BN: at BN = 64 each CTA stages 32 weight rows and one paired MMA (256 x 64 x 16) does
  // the work of two single-CTA instructions -- the 64-column data gradient (Toy_Net conv.3) was ISSUE-bound at one
  // 32-cycle MMA per instruction (tensor pipe 39 % active, profiles/ncu_gemm_v2.txt).
  static_assert(MC == 1 || (MC == 2 && !B_RES && (BN >= 128 || !B_MN)), "CTA pair: streamed weights; MN-major needs BN >= 128");
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work units: (group of MC consecutive row tiles) x column block; CTA `rank` of a cluster takes row tile
  // group * MC + rank.  A row tile past the end (odd tile count) is computed on zero-filled operands and masked.
  const int rank = MC > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / MC, unit_stride = gridDim.x / MC;
  const int num_units = ((num_m_tiles + MC - 1) / MC) * num_n_tiles;
  constexpr uint16_t kMask = (1u << MC) - 1;
  const int total_kb = args.num_taps * args.k_chunks;
  const int kStages = args.stages, kbs = args.kbs;
  const int a_stage_bytes = kbs * kAStageBytes;
  const int b_stage_bytes = kbs * kBStageBytes;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * a_stage_bytes;        // B_RES: total_kb resident tiles, else kStages ring slots
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (B_RES ? total_kb * kBStageBytes : kStages * b_stage_bytes));
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;               // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint64_t* b_ready = tmem_empty + 2;                      // [1] resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpilogueWarps * MC);      // pair: the leader collects both CTAs' epilogues
    }
    mbar_init(b_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (MC > 1) tmem_alloc_pair<kTmemCols>(tmem_slot); else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (MC > 1) cluster_sync_all(); else __syncthreads();   // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // operands come from the previous kernel of the stream: wait for it only now, after the setup
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {   // one elected lane: operands stay in uniform registers (no per-lane waterfall)
      const uint32_t tx_bytes = args.a_rows_bytes + (B_RES ? 0 : kBStageBytes);
      if (B_RES) {                                         // one-time: every (tap, chunk) weight tile, column block 0
        mbar_arrive_expect_tx(b_ready, static_cast<uint32_t>(total_kb) * kBStageBytes);
        for (int kb = 0; kb < total_kb; ++kb) {
          const int tap = kb / args.k_chunks, chunk = kb - tap * args.k_chunks;
          if (!B_MN) {
            tma_load_2d(smem_b + kb * kBStageBytes, &tmap_b, b_ready, kb * kBlockK, 0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(smem_b + kb * kBStageBytes + j * (kBlockK * 128), &tmap_b, b_ready,
                          tap * args.n_total + j * 64, chunk * kBlockK);
          }
        }
      }
      int ring_s = 0;
      uint32_t ring_phase = 0;
      for (int wk = 0;; ++wk) {
        const int unit = work_unit(args, unit0, unit_stride, num_units, wk);
        if (unit < 0) break;
        const int tile_n = unit % num_n_tiles, tile_m = (unit / num_n_tiles) * MC + rank;
        int tw, th, tn;
        decode_tile(args, tile_m, tw, th, tn);
        const TileTaps taps = unit_taps<MC>(args, (unit / num_n_tiles) * MC);
        const int ntx = taps.tx1 - taps.tx0 + 1;
        const int tile_kb = ntx * (taps.ty1 - taps.ty0 + 1) * args.k_chunks;     // K blocks of this tile
        const int steps_per_tile = (tile_kb + kbs - 1) / kbs;
        // A-box origin of tap (0, 0) and the weight column of this CTA; the (tap, channel chunk) of each K block is
        // advanced with counters.  ncu on the first version (profiles/ncu_fwd10_pair_v1.txt): the elected producer
        // lane spent ~85 % of its samples in two runtime integer divisions per K block (I2F / MUFU.RCP chains on one
        // in-order thread, ~600 clk) while its empty-slot wait never spun -- the producer, not L2 or shared memory,
        // capped every variant of this kernel at 70 % tensor-pipe activity.
        const int wb = tw * args.box_w + args.a_w_off, hb = th * args.box_h + args.a_h_off, n0 = tn * args.box_n;
        const int col0 = tile_n * BN + (MC > 1 ? rank * (BN / MC) : 0);
        int kb = 0, chunk = 0, tx = taps.tx0;
        int sx = args.tap_sign * taps.tx0, dh = args.tap_sign * taps.ty0;
        int tap = taps.ty0 * args.taps_w + taps.tx0;            // weight tap index: B K-offset tap*C, B_MN column tap*n_total
        for (int step = 0; step < steps_per_tile; ++step) {
          const int s = ring_s;
          const uint32_t phase = ring_phase;
          if (++ring_s == kStages) { ring_s = 0; ring_phase ^= 1; }
          mbar_wait(&empty_bar[s], phase ^ 1, 1);
          const int nkb = min(kbs, tile_kb - kb);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(nkb) * MC * tx_bytes);
          uint8_t* a_dst = smem_a + s * a_stage_bytes;
          uint8_t* b_dst = smem_b + s * b_stage_bytes;
          for (int j = 0; j < nkb; ++j, a_dst += kAStageBytes, b_dst += kBStageBytes) {
            if (MC > 1) tma_load_4d_pair(a_dst, &tmap_a, &full_bar[s], chunk * kBlockK, wb + sx, hb + dh, n0);
            else tma_load_4d(a_dst, &tmap_a, &full_bar[s], chunk * kBlockK, wb + sx, hb + dh, n0);
            if (B_RES) {
              // weights are resident
            } else if (!B_MN) {
              const int kcol = (tap * args.k_chunks + chunk) * kBlockK;
              if (MC > 1) tma_load_2d_pair(b_dst, &tmap_b, &full_bar[s], kcol, col0);
              else tma_load_2d(b_dst, &tmap_b, &full_bar[s], kcol, col0);
            } else {
              const int tap_col = tap * args.n_total + col0;
#pragma unroll
              for (int jj = 0; jj < BN / 64 / MC; ++jj) {
                if (MC > 1) tma_load_2d_pair(b_dst + jj * (kBlockK * 128), &tmap_b, &full_bar[s], tap_col + jj * 64, chunk * kBlockK);
                else tma_load_2d(b_dst + jj * (kBlockK * 128), &tmap_b, &full_bar[s], tap_col + jj * 64, chunk * kBlockK);
              }
            }
            ++kb;
            if (++chunk == args.k_chunks) {                 // next filter tap: one column over, or next filter row
              chunk = 0;
              ++tap;
              sx += args.tap_sign;
              if (++tx > taps.tx1) {
                tx = taps.tx0;
                sx = args.tap_sign * taps.tx0;
                dh += args.tap_sign;
                tap += args.taps_w - ntx;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if ((MC == 1 || rank == 0) && elect_one()) {   // one elected lane (of the leader CTA): uniform-register operands
      int ring_s = 0;
      uint32_t ring_phase = 0, t = 0;
      if (B_RES) mbar_wait(b_ready, 0, 5);
      for (int wk = 0;; ++wk, ++t) {
        const int unit = work_unit(args, unit0, unit_stride, num_units, wk);
        if (unit < 0) break;
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 4);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const TileTaps taps = unit_taps<MC>(args, (unit / num_n_tiles) * MC);
        const int ntx = taps.tx1 - taps.tx0 + 1;
        const int tile_kb = ntx * (taps.ty1 - taps.ty0 + 1) * args.k_chunks;
        const int steps_per_tile = (tile_kb + kbs - 1) / kbs;
        int chunk = 0, tx = taps.tx0, tap = taps.ty0 * args.taps_w + taps.tx0;   // resident weights: tile index of the K block
        for (int step = 0; step < steps_per_tile; ++step) {
          const int s = ring_s;
          const uint32_t phase = ring_phase;
          if (++ring_s == kStages) { ring_s = 0; ring_phase ^= 1; }
          mbar_wait(&full_bar[s], phase, 2);
          tc_fence_after();
          const int kb0 = step * kbs, nkb = min(kbs, tile_kb - kb0);
          for (int j = 0; j < nkb; ++j) {
            const int kb = kb0 + j;
            const uint32_t a_addr = smem_u32(smem_a + s * a_stage_bytes + j * kAStageBytes);
            const uint32_t b_addr = B_RES ? smem_u32(smem_b + (tap * args.k_chunks + chunk) * kBStageBytes)
                                          : smem_u32(smem_b + s * b_stage_bytes + j * kBStageBytes);
            if (B_RES) {
              if (++chunk == args.k_chunks) {
                chunk = 0;
                ++tap;
                if (++tx > taps.tx1) { tx = taps.tx0; tap += args.taps_w - ntx; }
              }
            }
            const uint64_t adesc0 = make_smem_desc(a_addr, 0, 1024);
            const uint64_t bdesc0 = B_MN ? make_smem_desc(b_addr, kBlockK * 128, 1024) : make_smem_desc(b_addr, 0, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t adesc = desc_advance(adesc0, k * (kUmmaK * 2));
              const uint64_t bdesc = desc_advance(bdesc0, k * (B_MN ? kUmmaK * 128 : kUmmaK * 2));
              if (MC > 1) mma_bf16_ss_pair(tmem_d, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
              else mma_bf16_ss(tmem_d, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if (MC > 1) mma_commit_pair(&empty_bar[s], kMask); else mma_commit(&empty_bar[s]);
        }
        if (MC > 1) mma_commit_pair(&tmem_full[acc], kMask); else mma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue: warps 2..9; quadrant = warp & 3, column half = (warp - 2) >> 2 =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int box_hw = args.box_h * args.box_w;
    const int dn = row / box_hw, rem = row - dn * box_hw;
    const int dh = rem / args.box_w, dw = rem - dh * args.box_w;
    constexpr int kChunksPerWarp = (BN / 32) / 2;
    // BatchNorm statistics: lane L accumulates column (chunk*32 + L) over every tile of this CTA.  The launcher
    // sizes the grid as a multiple of num_n_tiles, so a CTA always works on the same column block.
    float acc_s1[kChunksPerWarp], acc_s2[kChunksPerWarp];
#pragma unroll
    for (int cc = 0; cc < kChunksPerWarp; ++cc) acc_s1[cc] = acc_s2[cc] = 0.f;
    uint32_t t = 0;
    for (int wk = 0;; ++wk, ++t) {
      const int unit = work_unit(args, unit0, unit_stride, num_units, wk);
      if (unit < 0) break;
      const int tile_n = unit % num_n_tiles, tile_m = (unit / num_n_tiles) * MC + rank;
      int tw, th, tn;
      decode_tile(args, tile_m, tw, th, tn);
      const int n = tn * args.box_n + dn, h = th * args.box_h + dh, w = tw * args.box_w + dw;
      const bool valid = dn < args.box_n && n < args.out_n && h < args.out_h && w < args.out_w;
      const long long out_row = (static_cast<long long>(n) * args.out_h + h) * args.out_w + w;
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase, 3);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int c = half * kChunksPerWarp + cc;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const int col = tile_n * BN + c * 32;
        if (args.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(args.bias + col + j);
        }
        if (args.stat_partials != nullptr || args.stat_accum != nullptr) {
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
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (MC > 1) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
    }
    if (args.stat_accum != nullptr) {
      // fp64 accumulators shared by all CTAs (4 quadrants x 148 CTAs adds per column: negligible next to the tile loop)
      const int my_tile_n = unit0 % num_n_tiles;
#pragma unroll
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int col = my_tile_n * BN + (half * kChunksPerWarp + cc) * 32 + lane;
        atomicAdd(args.stat_accum + col, (double)acc_s1[cc]);
        atomicAdd(args.stat_accum + args.n_total + col, (double)acc_s2[cc]);
      }
    } else if (args.stat_partials != nullptr) {
      // one row group per CTA: [blockIdx.x][quad][2][n_total]; columns of other column blocks are written as 0
      const int my_tile_n = unit0 % num_n_tiles;
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

  // pair: a CTA may not retire (or free TMEM) while the leader's MMAs can still read its shared memory / write its
  // accumulator, or while the peer can still arrive on its barriers
  if (MC > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (MC > 1) tmem_dealloc_pair<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---- host-planned work lists (data gradient) ---------------------------------------------------------------------
// Cost of a unit = K blocks it actually runs (taps that reach the brick x channel chunks) with a floor for the
// epilogue; longest-processing-time-first onto the least loaded CTA (pair) that owns the unit's column block.  Plans are
// cached per problem shape in device memory (allocated outside CUDA-graph capture: the step engine's first, eager step).
struct PlanKey {
  int v[16];
  bool operator<(const PlanKey& o) const { return std::lexicographical_compare(v, v + 16, o.v, o.v + 16); }
};
struct Plan { int* dev; int len; };

template <int MC>
const Plan* plan_schedule(const KernelArgs& a, int num_units, int num_n_tiles, int num_ctas, cudaStream_t stream) {
  static std::map<PlanKey, Plan> cache;
  static std::mutex mu;
  PlanKey key{{a.tiles_w, a.tiles_h, a.box_w, a.box_h, a.box_n, a.out_w, a.out_h, a.out_n, a.taps_w, a.num_taps,
               a.k_chunks, a.a_w, a.a_h, a.pair_interleave, num_n_tiles * 4 + MC, num_ctas}};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
  std::vector<std::pair<int, int>> units(num_units);          // (cost, unit)
  for (int u = 0; u < num_units; ++u) {
    const TileTaps t = unit_taps<MC>(a, (u / num_n_tiles) * MC);
    const int kb = (t.tx1 - t.tx0 + 1) * (t.ty1 - t.ty0 + 1) * a.k_chunks;
    units[u] = {std::max(kb, 8) + 2, u};
  }
  std::stable_sort(units.begin(), units.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first > y.first; });
  std::vector<std::vector<int>> lists(num_ctas);
  std::vector<long long> load(num_ctas, 0);
  for (const auto& cu : units) {
    const int col = cu.second % num_n_tiles;
    int best = -1;
    for (int c = col; c < num_ctas; c += num_n_tiles)            // CTAs that own this column block (grid % num_n_tiles == 0)
      if (best < 0 || load[c] < load[best]) best = c;
    lists[best].push_back(cu.second);
    load[best] += cu.first;
  }
  size_t len = 1;
  for (const auto& l : lists) len = std::max(len, l.size() + 1);
  std::vector<int> table(num_ctas * len, -1);
  for (int c = 0; c < num_ctas; ++c) std::copy(lists[c].begin(), lists[c].end(), table.begin() + c * len);
  Plan plan{nullptr, (int)len};
  if (cudaMalloc(&plan.dev, table.size() * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  DDPT_CUDA_CHECK(cudaMemcpy(plan.dev, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice));
  return &cache.emplace(key, plan).first->second;
}

template <int BN, bool B_MN, bool B_RES, int MC = 1>
int launch_persistent_impl(const ConvGemmProblem& p, cudaStream_t stream) {
  constexpr int kBStageBytes = (BN / MC) * kBlockK * 2;
  const int total_kb_host = p.taps_h * p.taps_w * (p.a_c / kBlockK);
  // K blocks per pipeline stage: a barrier round trip (try_wait + fence + commit, ~300 cycles) should be covered by
  // >= ~512 tensor cycles of MMAs; one K block is 4 x (BN/4) cycles.  ncu on conv.3 dgrad (N = 64, one K block per
  // stage) showed the tensor pipe only 24 % active with every queue shallow (profiles/ncu_dgrad3.txt).
  int kbs = BN >= 256 ? 1 : (BN == 128 ? 2 : 3);
  if (kbs > total_kb_host) kbs = total_kb_host;
  const size_t budget = 227 * 1024 - 2048;
  const size_t resident_bytes = B_RES ? (size_t)total_kb_host * kBStageBytes : 0;
  const size_t stage_bytes = (size_t)kbs * (kAStageBytes + (B_RES ? 0 : kBStageBytes));
  int stages = (int)std::min<size_t>(8, (budget - resident_bytes) / stage_bytes);
  while (stages < 2 && kbs > 1) {           // keep at least two stages (>= 4 K blocks) in flight
    --kbs;
    stages = (int)std::min<size_t>(8, (budget - resident_bytes) / ((size_t)kbs * (kAStageBytes + (B_RES ? 0 : kBStageBytes))));
  }
  if (stages < 2) throw std::runtime_error("conv_gemm (persistent): tile does not fit shared memory");
  const size_t smem_bytes = 1024 + (size_t)stages * kbs * (kAStageBytes + (B_RES ? 0 : kBStageBytes)) + resident_bytes +
                            (2 * stages + 5) * 8 + 16;

  KernelArgs a{};
  a.stages = stages; a.kbs = kbs;
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
  a.a_w = p.a_w; a.a_h = p.a_h; a.fp16 = p.fp16;
  a.pair_interleave = (MC == 2 && tiles_n % 2 == 0 && p.taps_h * p.taps_w > 1) ? 1 : 0;

  if (p.a_c % kBlockK != 0 || p.n_total % BN != 0 || p.box_w * p.box_h * p.box_n > kBlockM)
    throw std::runtime_error("conv_gemm: unsupported shape (C % 64, N % BLOCK_N, brick <= 128 rows)");
  if (p.split_k > 1) throw std::runtime_error("conv_gemm (persistent): split_k is handled by the non-persistent kernel");

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
    const uint32_t b_box[2] = {(uint32_t)kBlockK, (uint32_t)(BN / MC)};
    tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);
  } else {
    const uint64_t b_dims[2] = {(uint64_t)a.num_taps * p.n_total, (uint64_t)p.a_c};
    const uint64_t b_strides[1] = {(uint64_t)p.b_ld * 2};
    const uint32_t b_box[2] = {64u, (uint32_t)kBlockK};
    tmap_b = make_tensor_map_bf16(p.b, 2, b_dims, b_strides, b_box);
  }

  auto kernel = conv_gemm_persistent_kernel<BN, B_MN, B_RES, MC>;
  static bool configured = false;
  static int max_units = 0;                    // co-resident CTAs (MC = 1) or clusters (MC = 2) at one per SM
  const int tiles_m = a.tiles_w * a.tiles_h * tiles_n;
  const int tiles_n_cols = p.n_total / BN;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.numAttrs = 1;
  if (MC > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = MC; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  if (!configured) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    max_units = sm_count();
    if (MC > 1) {
      // GPCs with an odd number of usable SMs cannot host a last pair: ask the driver how many clusters co-reside
      cfg.gridDim = dim3(sm_count() / MC * MC);
      cfg.dynamicSmemBytes = 227 * 1024;
      DDPT_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_units, kernel, &cfg));
      cfg.dynamicSmemBytes = smem_bytes;
      if (max_units < 1) throw std::runtime_error("conv_gemm (persistent): no co-resident cluster");
      if (std::getenv("DDPT_GEMM_DEBUG") != nullptr)
        std::fprintf(stderr, "ddpt conv_gemm<%d,%d,mc%d>: %d co-resident clusters on %d SMs\n", BN, (int)B_MN, MC,
                     max_units, sm_count());
    }
    configured = true;
  }
  const int total_units = ((tiles_m + MC - 1) / MC) * tiles_n_cols;
  int units = total_units < max_units ? total_units : max_units;
  units -= units % tiles_n_cols;               // a CTA keeps one column block (statistics accumulate in registers)
  if (units < tiles_n_cols) units = tiles_n_cols;
  const int grid = units * MC;
  cfg.gridDim = dim3(grid);
  if (p.tap_sign < 0 && a.num_taps > 1 && total_units > units && std::getenv("DDPT_NO_TILE_PLAN") == nullptr) {
    // uneven tiles (border bricks skip filter taps) and more than one unit per CTA: hand every CTA a planned list
    const Plan* plan = plan_schedule<MC>(a, total_units, tiles_n_cols, units, stream);
    if (plan != nullptr) { a.schedule = plan->dev; a.sched_len = plan->len; }
  }
  DDPT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmap_a, tmap_b, a, tiles_m, tiles_n_cols));
  return p.stat_partials != nullptr ? grid : tiles_m;   // statistic row groups (x4 rows) written
}

}  // namespace

int launch_conv_gemm_persistent(const ConvGemmProblem& p, int block_n, cudaStream_t stream, int cta_pair) {
  // resident weights: one column block, and taps*C*BLOCK_N*2 bytes + a 4-deep activation ring fit in shared memory
  const long long b_bytes = (long long)p.taps_h * p.taps_w * p.a_c * block_n * 2;
  const bool resident = p.n_total == block_n && b_bytes + 4 * kAStageBytes + 2048 <= 227 * 1024 && b_bytes < (1 << 20);
  if (cta_pair && block_n == 64 && !p.b_mn_major) return launch_persistent_impl<64, false, false, 2>(p, stream);
  if (cta_pair && !resident && block_n >= 128) {          // CTA pairs (cta_group::2): 256-row tiles, weights split
    if (p.b_mn_major) {
      if (block_n == 128) return launch_persistent_impl<128, true, false, 2>(p, stream);
      if (block_n == 256) return launch_persistent_impl<256, true, false, 2>(p, stream);
    } else {
      if (block_n == 128) return launch_persistent_impl<128, false, false, 2>(p, stream);
      if (block_n == 256) return launch_persistent_impl<256, false, false, 2>(p, stream);
    }
  }
  if (p.b_mn_major) {
    switch (block_n) {
      case 64: return resident ? launch_persistent_impl<64, true, true>(p, stream) : launch_persistent_impl<64, true, false>(p, stream);
      case 128: return resident ? launch_persistent_impl<128, true, true>(p, stream) : launch_persistent_impl<128, true, false>(p, stream);
      case 256: return launch_persistent_impl<256, true, false>(p, stream);
    }
  } else {
    switch (block_n) {
      case 64: return resident ? launch_persistent_impl<64, false, true>(p, stream) : launch_persistent_impl<64, false, false>(p, stream);
      case 128: return resident ? launch_persistent_impl<128, false, true>(p, stream) : launch_persistent_impl<128, false, false>(p, stream);
      case 256: return launch_persistent_impl<256, false, false>(p, stream);
    }
  }
  throw std::runtime_error("conv_gemm: block_n must be 64, 128 or 256");
}

}  // namespace ddpt
