// Block-scaled fp8 (mxf8f6f4) descriptor / scale-factor layout probe: one CTA, one 128 x 64 x 128 tile,
// tcgen05.mma.kind::mxf8f6f4.block_scale with the UE8M0 scale factors written to TMEM by tcgen05.st from
// host-supplied words, so that the (lane, column, byte) -> (row, k-block) mapping can be read off the result.
// Tooling for benchmarks/probe_umma_mx.py; not on a hot path.
#include <cuda.h>
#include <stdexcept>
#include "ddpt_common.cuh"
#include "umma.cuh"
#include "nn/tensor_map.h"

namespace ddpt {

struct MxProbeArgs {
  const uint32_t* sfa_words;   // [128 lanes][4 columns]
  const uint32_t* sfb_words;   // [128 lanes][4 columns]
  int sfa_col, sfb_col;        // TMEM column offsets (relative to the allocation) of the scale factors
  int sf_id_mode;              // 0: sf ids 0 for all MMAs, 1: sf id = k (byte k of the word)
  int ksteps;                  // number of K=32 MMAs (<= 4)
  float* out;                  // [128][64]
};

namespace {
using namespace umma;

__device__ __forceinline__ void tmem_st_32x32_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// kind::mxf8f6f4 block-scaled instruction descriptor: e4m3 x e4m3 -> f32, UE8M0 scales, K-major operands
__device__ __forceinline__ uint32_t make_idesc_mxf8(uint32_t m, uint32_t n, uint32_t a_sf_id, uint32_t b_sf_id) {
  return ((b_sf_id & 3u) << 4)       // b_sf_id
         | (0u << 7) | (0u << 10)     // a/b format: E4M3
         | ((n >> 3) << 17)
         | (1u << 23)                 // scale format: UE8M0
         | ((m >> 4) << 24)
         | ((a_sf_id & 3u) << 29);
}

__device__ __forceinline__ void mma_mxf8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
umma_mx_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const MxProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                 // 128 rows x 128 B (K = 128 fp8)
  uint8_t* smem_b = smem + 16 * 1024;     // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // scale factors: thread t owns TMEM lane t; 4 consecutive columns each for A and B
  {
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t* wa = p.sfa_words + threadIdx.x * 4;
    const uint32_t* wb = p.sfb_words + threadIdx.x * 4;
    tmem_st_32x32_x4(lane_addr + p.sfa_col, wa[0], wa[1], wa[2], wa[3]);
    tmem_st_32x32_x4(lane_addr + p.sfb_col, wb[0], wb[1], wb[2], wb[3]);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], (128 + 64) * 128);
    tma_load_2d(smem_a, &tmap_a, &bars[0], 0, 0);
    tma_load_2d(smem_b, &tmap_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0, 51);
    tc_fence_after();
    for (int k = 0; k < p.ksteps; ++k) {
      const uint32_t sf = p.sf_id_mode ? k : 0;
      const uint64_t ad = make_smem_desc(smem_u32(smem_a) + k * 32, 0, 1024);
      const uint64_t bd = make_smem_desc(smem_u32(smem_b) + k * 32, 0, 1024);
      mma_mxf8_ss(tmem_base, ad, bd, make_idesc_mxf8(128, 64, sf, sf), tmem_base + p.sfa_col, tmem_base + p.sfb_col,
                  k != 0 ? 1u : 0u);
    }
    mma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0, 52);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    float* o = p.out + (warp * 32 + lane) * 64 + c * 32;
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}
}  // namespace

void launch_umma_mx_probe(const void* a_fp8 /*[128][128]*/, const void* b_fp8 /*[64][128]*/, const MxProbeArgs& p,
                          cudaStream_t stream) {
  const uint64_t a_dims[2] = {128, 128}, b_dims[2] = {128, 64};
  const uint64_t str[1] = {128};
  const uint32_t a_box[2] = {128, 128}, b_box[2] = {128, 64};
  CUtensorMap ta = make_tensor_map_u8(a_fp8, 2, a_dims, str, a_box);
  CUtensorMap tb = make_tensor_map_u8(b_fp8, 2, b_dims, str, b_box);
  const size_t smem = 1024 + 32 * 1024 + 64;
  umma_mx_probe_kernel<<<1, 128, smem, stream>>>(ta, tb, p);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
