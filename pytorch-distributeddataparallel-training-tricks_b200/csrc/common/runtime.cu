// Process-wide runtime helpers (error reporting, device properties).
#include <cstdlib>
#include "../include/ddpt_common.cuh"
#include <cstdio>
#include <stdexcept>
#include <string>

namespace ddpt {

void report_cuda_error(cudaError_t e, const char* expr, const char* file, int line) {
  std::string msg = std::string("ddpt_b200 CUDA error: ") + cudaGetErrorString(e) + " at " + file + ":" +
                    std::to_string(line) + " (" + expr + ")";
  throw std::runtime_error(msg);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  DDPT_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  DDPT_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 64) cached[dev] = n;
  return n;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = std::getenv("DDPT_PDL");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached != 0;
}

}  // namespace ddpt
