#include "tensor_map.h"
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <string>

namespace ddpt {
namespace {

PFN_cuTensorMapEncodeTiled_v12000 resolve_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || sym == nullptr)
      throw std::runtime_error("ddpt_b200: cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  });
  return fn;
}

}  // namespace

static CUtensorMap make_tensor_map_typed(CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                                         const uint64_t* strides_bytes, const uint32_t* box) {
  CUtensorMap map;
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bdim[5];
  cuuint32_t estride[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estride[i] = 1;
    if (i > 0) gstride[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = resolve_encode()(&map, dtype, (cuuint32_t)rank, const_cast<void*>(base),
                                gdim, gstride, bdim, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string msg = "ddpt_b200: cuTensorMapEncodeTiled failed (CUresult " + std::to_string((int)r) + "), rank " +
                      std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string(dims[i]) + "/box" + std::to_string(box[i]);
    throw std::runtime_error(msg);
  }
  return map;
}

CUtensorMap make_tensor_map_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box) {
  return make_tensor_map_typed(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}

CUtensorMap make_tensor_map_u8(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                               const uint32_t* box) {
  return make_tensor_map_typed(CU_TENSOR_MAP_DATA_TYPE_UINT8, base, rank, dims, strides_bytes, box);
}

}  // namespace ddpt
