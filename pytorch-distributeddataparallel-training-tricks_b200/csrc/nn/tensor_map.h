// Host-side TMA descriptor construction (cuTensorMapEncodeTiled via the runtime's driver entry
// point, so the extension does not link libcuda directly).
#pragma once
#include <cuda.h>
#include <cstdint>
#include <stdexcept>

namespace ddpt {

// bf16 tensor, `rank` dims (innermost first), strides in BYTES for dims 1..rank-1, 128-byte swizzle,
// out-of-bounds elements read as zero.
CUtensorMap make_tensor_map_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box);
// Same for 1-byte elements (fp8 operands of the block-scaled path).
CUtensorMap make_tensor_map_u8(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                               const uint32_t* box);

}  // namespace ddpt
