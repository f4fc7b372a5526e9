// Bindings of the hardware probes (benchmarks/probes/*.cu): NOT part of the shipped extension.  Built on demand by
// benchmarks/probes/build.py into benchmarks/probes/_ddpt_probes*.so and used by benchmarks/probe_umma_*.py.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

namespace ddpt {
struct UmmaProbeArgs {
  int a_rows, a_blocks, b_rows, b_blocks;
  int a_start_off, a_lbo, a_sbo, a_base_offset, a_mn_major, a_kstep_bytes;
  int b_start_off, b_lbo, b_sbo, b_base_offset, b_mn_major, b_kstep_bytes;
  int ksteps;
  float* out;
};
void launch_umma_probe(const void* a, int a_ld, int a_total_rows, const void* b, int b_ld, int b_total_rows,
                       const UmmaProbeArgs& p, cudaStream_t stream);
}  // namespace ddpt

// a: [rows, a_ld] bf16, b: [rows, b_ld] bf16; cfg = 17 ints in UmmaProbeArgs order.  Returns D [128, 64] fp32.
static at::Tensor umma_probe(at::Tensor a, at::Tensor b, std::vector<int64_t> cfg) {
  TORCH_CHECK(cfg.size() == 17 && a.is_cuda() && b.is_cuda() && a.dim() == 2 && b.dim() == 2);
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && a.is_contiguous() && b.is_contiguous());
  c10::cuda::CUDAGuard guard(a.device());
  at::Tensor out = at::zeros({128, 64}, a.options().dtype(at::kFloat));
  ddpt::UmmaProbeArgs p{(int)cfg[0], (int)cfg[1], (int)cfg[2], (int)cfg[3], (int)cfg[4], (int)cfg[5], (int)cfg[6],
                        (int)cfg[7], (int)cfg[8], (int)cfg[9], (int)cfg[10], (int)cfg[11], (int)cfg[12], (int)cfg[13],
                        (int)cfg[14], (int)cfg[15], (int)cfg[16], out.data_ptr<float>()};
  ddpt::launch_umma_probe(a.data_ptr(), (int)a.size(1), (int)a.size(0), b.data_ptr(), (int)b.size(1), (int)b.size(0), p,
                          at::cuda::getCurrentCUDAStream().stream());
  return out;
}

namespace ddpt {
struct MxProbeArgs {
  const uint32_t* sfa_words; const uint32_t* sfb_words;
  int sfa_col, sfb_col, sf_id_mode, ksteps;
  float* out;
};
void launch_umma_mx_probe(const void* a_fp8, const void* b_fp8, const MxProbeArgs& p, cudaStream_t stream);
}  // namespace ddpt

static at::Tensor umma_mx_probe(at::Tensor a, at::Tensor b, at::Tensor sfa, at::Tensor sfb, int64_t sfa_col,
                                int64_t sfb_col, int64_t sf_id_mode, int64_t ksteps) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kByte && a.numel() == 128 * 128 && a.is_contiguous());
  TORCH_CHECK(b.is_cuda() && b.scalar_type() == at::kByte && b.numel() == 64 * 128 && b.is_contiguous());
  TORCH_CHECK(sfa.scalar_type() == at::kInt && sfa.numel() == 512 && sfb.scalar_type() == at::kInt && sfb.numel() == 512);
  c10::cuda::CUDAGuard guard(a.device());
  at::Tensor out = at::zeros({128, 64}, a.options().dtype(at::kFloat));
  ddpt::MxProbeArgs p{reinterpret_cast<const uint32_t*>(sfa.data_ptr<int>()),
                      reinterpret_cast<const uint32_t*>(sfb.data_ptr<int>()), (int)sfa_col, (int)sfb_col,
                      (int)sf_id_mode, (int)ksteps, out.data_ptr<float>()};
  ddpt::launch_umma_mx_probe(a.data_ptr(), b.data_ptr(), p, at::cuda::getCurrentCUDAStream().stream());
  return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "ddpt_b200 hardware probes (tcgen05 descriptor addressing, MXFP8 scale-factor layout)";
  m.def("umma_probe", &umma_probe);
  m.def("umma_mx_probe", &umma_mx_probe);
}
