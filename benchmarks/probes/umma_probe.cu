// Descriptor probe: one CTA, one 128 x 64 x (16*ksteps) tcgen05.mma chain with every shared-memory descriptor
// field supplied by the host.  Used by tests/benchmarks to pin down how the tensor core addresses swizzled
// operands whose start is not 1024-byte aligned (needed for halo-reuse convolution tiles) -- see
// benchmarks/probe_umma_desc.py.  Not on any hot path.
#include <cuda.h>
#include <stdexcept>
#include "ddpt_common.cuh"
#include "umma.cuh"
#include "nn/tensor_map.h"

namespace ddpt {

struct UmmaProbeArgs {
  int a_rows, a_blocks;        // TMA: a_blocks boxes of [a_rows x 64 bf16], block j from columns [64j, 64j+64)
  int b_rows, b_blocks;
  int a_start_off, a_lbo, a_sbo, a_base_offset, a_mn_major, a_kstep_bytes;
  int b_start_off, b_lbo, b_sbo, b_base_offset, b_mn_major, b_kstep_bytes;
  int ksteps;                  // number of K=16 MMAs
  float* out;                  // [128][64] fp32
};

namespace {
using namespace umma;

__device__ __forceinline__ uint64_t desc_with_base(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_offset) {
  return make_smem_desc(addr, lbo, sbo) | (static_cast<uint64_t>(base_offset & 7u) << 49);
}

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const UmmaProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + 64 * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 128 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t bytes = (p.a_rows * p.a_blocks + p.b_rows * p.b_blocks) * 128;
    mbar_arrive_expect_tx(&bars[0], bytes);
    for (int j = 0; j < p.a_blocks; ++j) tma_load_2d(smem_a + j * p.a_rows * 128, &tmap_a, &bars[0], 64 * j, 0);
    for (int j = 0; j < p.b_blocks; ++j) tma_load_2d(smem_b + j * p.b_rows * 128, &tmap_b, &bars[0], 64 * j, 0);
    mbar_wait(&bars[0], 0, 21);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, p.a_mn_major != 0, p.b_mn_major != 0);
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t ad = desc_with_base(smem_u32(smem_a) + p.a_start_off + k * p.a_kstep_bytes, p.a_lbo, p.a_sbo,
                                         p.a_base_offset);
      const uint64_t bd = desc_with_base(smem_u32(smem_b) + p.b_start_off + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo,
                                         p.b_base_offset);
      mma_bf16_ss(tmem_base, ad, bd, idesc, k != 0 ? 1u : 0u);
    }
    mma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0, 22);
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
    tmem_dealloc<64>(tmem_base);
  }
}
}  // namespace

void launch_umma_probe(const void* a, int a_ld, int a_total_rows, const void* b, int b_ld, int b_total_rows,
                       const UmmaProbeArgs& p, cudaStream_t stream) {
  if (p.a_rows * p.a_blocks * 128 > 64 * 1024 || p.b_rows * p.b_blocks * 128 > 64 * 1024 || p.a_rows > 256 || p.b_rows > 256)
    throw std::runtime_error("umma_probe: operand does not fit the 64 KiB staging area / 256-row box");
  const uint64_t a_dims[2] = {(uint64_t)a_ld, (uint64_t)a_total_rows};
  const uint64_t a_str[1] = {(uint64_t)a_ld * 2};
  const uint32_t a_box[2] = {64u, (uint32_t)p.a_rows};
  const uint64_t b_dims[2] = {(uint64_t)b_ld, (uint64_t)b_total_rows};
  const uint64_t b_str[1] = {(uint64_t)b_ld * 2};
  const uint32_t b_box[2] = {64u, (uint32_t)p.b_rows};
  CUtensorMap ta = make_tensor_map_bf16(a, 2, a_dims, a_str, a_box);
  CUtensorMap tb = make_tensor_map_bf16(b, 2, b_dims, b_str, b_box);
  const size_t smem = 1024 + 128 * 1024 + 64;
  static bool configured = false;
  if (!configured) {
    DDPT_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  umma_probe_kernel<<<1, 128, smem, stream>>>(ta, tb, p);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
