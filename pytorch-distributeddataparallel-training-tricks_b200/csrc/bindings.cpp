// Python bindings (pybind11 / torch extension) for the ddpt_b200 sm_100a kernels.
// Tensors are validated here; kernels receive raw pointers and the current CUDA stream.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "include/ddpt_amp.h"
#include "include/ddpt_nn.h"
#include "include/ddpt_optim.h"
#include "bindings_decl.h"

namespace {

cudaStream_t current_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check_cuda(const at::Tensor& t, const char* name, at::ScalarType dtype) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == dtype, name, " has dtype ", t.scalar_type(), ", expected ", dtype);
}

// ---- amp --------------------------------------------------------------------------------------
ddpt::TensorList make_list(const std::vector<at::Tensor>& in0, const std::vector<at::Tensor>* in1,
                           const std::vector<at::Tensor>& out, size_t begin, size_t end) {
  ddpt::TensorList l{};
  l.count = static_cast<int>(end - begin);
  for (size_t i = begin; i < end; ++i) {
    const size_t k = i - begin;
    TORCH_CHECK(in0[i].is_non_overlapping_and_dense(), "multi_tensor: dense tensors only");
    l.numel[k] = in0[i].numel();
    l.in0[k] = in0[i].data_ptr();
    l.in1[k] = in1 ? (*in1)[i].data_ptr() : nullptr;
    l.out[k] = out[i].data_ptr();
  }
  return l;
}

void multi_tensor_scale(std::vector<at::Tensor> in, std::vector<at::Tensor> out, at::Tensor scale,
                        c10::optional<at::Tensor> found_inf) {
  TORCH_CHECK(in.size() == out.size(), "multi_tensor_scale: list length mismatch");
  if (in.empty()) return;
  c10::cuda::CUDAGuard guard(in[0].device());
  check_cuda(scale, "scale", at::kFloat);
  for (size_t i = 0; i < in.size(); ++i) {
    check_cuda(in[i], "in", at::kFloat);
    check_cuda(out[i], "out", at::kFloat);
    TORCH_CHECK(in[i].numel() == out[i].numel(), "multi_tensor_scale: size mismatch");
  }
  float* flag = found_inf.has_value() ? found_inf->data_ptr<float>() : nullptr;
  for (size_t b = 0; b < in.size(); b += ddpt::kMaxTensorsPerLaunch) {
    size_t e = std::min(in.size(), b + ddpt::kMaxTensorsPerLaunch);
    ddpt::launch_multi_tensor_scale(make_list(in, nullptr, out, b, e), scale.data_ptr<float>(), flag,
                                    current_stream());
  }
}

void multi_tensor_axpby(std::vector<at::Tensor> x, std::vector<at::Tensor> y, std::vector<at::Tensor> out,
                        double a, double b, int64_t arg_to_check, c10::optional<at::Tensor> found_inf) {
  TORCH_CHECK(x.size() == y.size() && x.size() == out.size(), "multi_tensor_axpby: list length mismatch");
  if (x.empty()) return;
  c10::cuda::CUDAGuard guard(x[0].device());
  for (size_t i = 0; i < x.size(); ++i) {
    check_cuda(x[i], "x", at::kFloat); check_cuda(y[i], "y", at::kFloat); check_cuda(out[i], "out", at::kFloat);
  }
  float* flag = found_inf.has_value() ? found_inf->data_ptr<float>() : nullptr;
  for (size_t s = 0; s < x.size(); s += ddpt::kMaxTensorsPerLaunch) {
    size_t e = std::min(x.size(), s + ddpt::kMaxTensorsPerLaunch);
    ddpt::launch_multi_tensor_axpby(make_list(x, &y, out, s, e), (float)a, (float)b, (int)arg_to_check, flag,
                                    current_stream());
  }
}

void unscale_arena(at::Tensor grads, at::Tensor scaler_state) {
  check_cuda(grads, "grads", at::kFloat); check_cuda(scaler_state, "scaler_state", at::kFloat);
  TORCH_CHECK(grads.is_contiguous() && scaler_state.numel() >= 5);
  c10::cuda::CUDAGuard guard(grads.device());
  ddpt::launch_unscale_arena(grads.data_ptr<float>(), grads.numel(), scaler_state.data_ptr<float>(), current_stream());
}

void update_scale(at::Tensor scaler_state, int64_t growth_interval, double min_scale, double max_scale) {
  check_cuda(scaler_state, "scaler_state", at::kFloat);
  c10::cuda::CUDAGuard guard(scaler_state.device());
  ddpt::launch_update_scale(scaler_state.data_ptr<float>(), (int)growth_interval, (float)min_scale, (float)max_scale,
                            current_stream());
}

// ---- optimizer ----------------------------------------------------------------------------------
void fused_lookahead_sgd(at::Tensor params, at::Tensor grads, at::Tensor momentum, c10::optional<at::Tensor> slow,
                         c10::optional<at::Tensor> shadow, at::Tensor hyper, at::Tensor state,
                         c10::optional<at::Tensor> found_inf, bool nesterov, bool zero_grads) {
  check_cuda(params, "params", at::kFloat); check_cuda(grads, "grads", at::kFloat);
  check_cuda(momentum, "momentum", at::kFloat); check_cuda(hyper, "hyper", at::kFloat);
  check_cuda(state, "state", at::kInt);
  TORCH_CHECK(params.is_contiguous() && grads.is_contiguous() && momentum.is_contiguous());
  TORCH_CHECK(params.numel() % 4 == 0 && grads.numel() == params.numel() && momentum.numel() == params.numel(),
              "fused_lookahead_sgd: arenas must have equal length, multiple of 4");
  TORCH_CHECK(hyper.numel() >= ddpt::kHyperSlots && state.numel() >= ddpt::kOptStateSlots);
  c10::cuda::CUDAGuard guard(params.device());
  ddpt::FusedSgdArgs a{};
  a.params = params.data_ptr<float>();
  a.grads = grads.data_ptr<float>();
  a.momentum = momentum.data_ptr<float>();
  a.slow = slow.has_value() ? slow->data_ptr<float>() : nullptr;
  if (shadow.has_value()) {
    TORCH_CHECK(shadow->is_cuda() && (shadow->scalar_type() == at::kBFloat16 || shadow->scalar_type() == at::kHalf),
                "shadow must be a CUDA bf16 or fp16 tensor");
    TORCH_CHECK(shadow->numel() == params.numel());
    a.shadow = reinterpret_cast<__nv_bfloat16*>(shadow->data_ptr());
    a.shadow_fp16 = shadow->scalar_type() == at::kHalf ? 1 : 0;
  }
  a.grads_to_zero = zero_grads ? grads.data_ptr<float>() : nullptr;
  a.hyper = hyper.data_ptr<float>();
  a.state = state.data_ptr<int>();
  a.found_inf = found_inf.has_value() ? found_inf->data_ptr<float>() : nullptr;
  a.numel = params.numel();
  a.nesterov = nesterov ? 1 : 0;
  ddpt::launch_fused_lookahead_sgd(a, current_stream());
}

void cast_bf16(at::Tensor src, at::Tensor dst) {
  check_cuda(src, "src", at::kFloat);
  TORCH_CHECK(dst.is_cuda() && (dst.scalar_type() == at::kBFloat16 || dst.scalar_type() == at::kHalf));
  TORCH_CHECK(src.is_contiguous() && dst.is_contiguous() && src.numel() == dst.numel() && src.numel() % 4 == 0);
  c10::cuda::CUDAGuard guard(src.device());
  ddpt::launch_cast_bf16(src.data_ptr<float>(), reinterpret_cast<__nv_bfloat16*>(dst.data_ptr()), src.numel(),
                         current_stream(), dst.scalar_type() == at::kHalf ? 1 : 0);
}

// ---- tensor-core GEMM / conv ----------------------------------------------------------------------
// a: [N, H, W, C] bf16 contiguous.  b: 2-D bf16.  out: [out_n*out_h*out_w, ld_out] bf16/fp32.
int64_t conv_gemm(at::Tensor a, at::Tensor b, at::Tensor out, std::vector<int64_t> out_nhw,
                  std::vector<int64_t> box_nhw, std::vector<int64_t> taps_hw, int64_t tap_sign,
                  std::vector<int64_t> a_off_hw, bool b_mn_major, c10::optional<at::Tensor> bias, bool relu,
                  c10::optional<at::Tensor> stat_partials, int64_t split_k, int64_t block_n, int64_t mode,
                  bool split_slabs) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && (a.scalar_type() == at::kBFloat16 || a.scalar_type() == at::kHalf) &&
              b.scalar_type() == a.scalar_type(), "a and b must be CUDA tensors of the same 16-bit type (bf16 or fp16)");
  TORCH_CHECK(a.dim() == 4 && a.is_contiguous(), "a must be contiguous [N,H,W,C]");
  TORCH_CHECK(b.dim() == 2 && b.stride(1) == 1, "b must be 2-D with unit inner stride");
  TORCH_CHECK(out.dim() == 2 && out.stride(1) == 1 && (out.scalar_type() == a.scalar_type() || out.scalar_type() == at::kFloat),
              "out must be fp32 or the operands' 16-bit type");
  c10::cuda::CUDAGuard guard(a.device());
  ddpt::ConvGemmProblem p{};
  p.a = reinterpret_cast<const __nv_bfloat16*>(a.data_ptr());
  p.a_n = (int)a.size(0); p.a_h = (int)a.size(1); p.a_w = (int)a.size(2); p.a_c = (int)a.size(3);
  p.out_n = (int)out_nhw[0]; p.out_h = (int)out_nhw[1]; p.out_w = (int)out_nhw[2];
  p.box_n = (int)box_nhw[0]; p.box_h = (int)box_nhw[1]; p.box_w = (int)box_nhw[2];
  p.taps_h = (int)taps_hw[0]; p.taps_w = (int)taps_hw[1]; p.tap_sign = (int)tap_sign;
  p.a_h_off = (int)a_off_hw[0]; p.a_w_off = (int)a_off_hw[1];
  p.b = reinterpret_cast<const __nv_bfloat16*>(b.data_ptr());
  p.b_mn_major = b_mn_major ? 1 : 0;
  p.b_ld = b.stride(0);
  p.fp16 = a.scalar_type() == at::kHalf ? 1 : 0;
  const int64_t taps = p.taps_h * p.taps_w;
  if (b_mn_major) {
    TORCH_CHECK(b.size(0) == p.a_c && b.size(1) % taps == 0, "MN-major b must be [C_k, taps*N]");
    p.n_total = (int)(b.size(1) / taps);
  } else {
    TORCH_CHECK(b.size(1) == taps * p.a_c, "K-major b must be [N, taps*C]");
    p.n_total = (int)b.size(0);
  }
  TORCH_CHECK(out.size(0) >= (int64_t)p.out_n * p.out_h * p.out_w && out.size(1) >= p.n_total, "out too small");
  p.out = out.data_ptr(); p.ld_out = out.stride(0); p.out_fp32 = out.scalar_type() == at::kFloat;
  if (bias.has_value()) { check_cuda(*bias, "bias", at::kFloat); TORCH_CHECK(bias->numel() >= p.n_total); p.bias = bias->data_ptr<float>(); }
  p.relu = relu ? 1 : 0;
  if (stat_partials.has_value()) {
    if (stat_partials->scalar_type() == at::kDouble) {     // fp64 accumulators [2][N]: epilogues add, no partial rows
      check_cuda(*stat_partials, "stat_partials", at::kDouble);
      TORCH_CHECK(stat_partials->numel() >= 2 * p.n_total && split_k <= 1, "stat accumulators: need [2][N] doubles, no split-K");
      p.stat_accum = stat_partials->data_ptr<double>();
    } else {
      check_cuda(*stat_partials, "stat_partials", at::kFloat);
      p.stat_partials = stat_partials->data_ptr<float>();
    }
  }
  p.split_k = (int)split_k;
  if (split_slabs && split_k > 1) {       // out is [split_k * rows, ld]: one fp32 slab per split, plain stores
    const int64_t rows = (int64_t)p.out_n * p.out_h * p.out_w;
    TORCH_CHECK(p.out_fp32 && out.size(0) >= split_k * rows && mode == 0, "split slabs: fp32 [split_k * rows, N] output, per-tile kernel");
    p.split_slab_stride = rows * p.ld_out;
  }
  if (p.stat_partials != nullptr) {
    const int64_t tiles = (int64_t)((p.out_w + p.box_w - 1) / p.box_w) * ((p.out_h + p.box_h - 1) / p.box_h) *
                          ((p.out_n + p.box_n - 1) / p.box_n);
    // persistent kernels write one row group per CTA (grid <= #SMs); the per-tile kernel one per M tile
    const int64_t n_tiles = p.n_total / block_n;
    const int64_t sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    const int64_t groups = (mode >= 1 && split_k <= 1) ? std::max<int64_t>(n_tiles, std::min<int64_t>(tiles * n_tiles, sms)) : tiles;
    TORCH_CHECK(stat_partials->numel() >= groups * 8 * p.n_total, "stat_partials too small: need ", groups * 8 * p.n_total);
    TORCH_CHECK(split_k <= 1, "statistics need the full reduction in one CTA");
  }
  // mode 4: 3x3 data gradient with 64 input channels, filter columns merged into N (conv_dgrad_kwmerge.cu)
  if (mode == 4) return ddpt::launch_conv_dgrad_kwmerge(p, current_stream());
  // mode: 0 = one CTA per tile (split-K capable), 1 = persistent, 2 = persistent + halo reuse
  if (mode == 2 && split_k <= 1) return ddpt::launch_conv_gemm_halo(p, (int)block_n, current_stream());
  // mode 3: persistent kernel as CTA pairs (tcgen05 cta_group::2, 256-row tiles)
  if (mode >= 1 && split_k <= 1) return ddpt::launch_conv_gemm_persistent(p, (int)block_n, current_stream(), mode == 3);
  return ddpt::launch_conv_gemm(p, (int)block_n, current_stream());
}

// dy: [N,Ho,Wo,M] bf16, x: [N,H,W,C] bf16, dw: [M, taps_h, taps_w, C] fp32 (dense).
void wgrad_gemm(at::Tensor dy, at::Tensor x, at::Tensor dw, std::vector<int64_t> taps_hw,
                std::vector<int64_t> box_nhw, int64_t split_k, bool accumulate, int64_t block_n, bool multitap,
                c10::optional<at::Tensor> slab) {
  TORCH_CHECK(dy.is_cuda() && x.is_cuda() && (dy.scalar_type() == at::kBFloat16 || dy.scalar_type() == at::kHalf) &&
              x.scalar_type() == dy.scalar_type(), "dy and x must be CUDA tensors of the same 16-bit type");
  check_cuda(dw, "dw", at::kFloat);
  TORCH_CHECK(dy.dim() == 4 && dy.is_contiguous() && x.dim() == 4 && x.is_contiguous() && dw.is_contiguous());
  TORCH_CHECK(dy.size(0) == x.size(0));
  c10::cuda::CUDAGuard guard(dy.device());
  ddpt::WgradProblem p{};
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy.data_ptr());
  p.out_n = (int)dy.size(0); p.out_h = (int)dy.size(1); p.out_w = (int)dy.size(2); p.m_total = (int)dy.size(3);
  p.x = reinterpret_cast<const __nv_bfloat16*>(x.data_ptr());
  p.in_h = (int)x.size(1); p.in_w = (int)x.size(2); p.c_total = (int)x.size(3);
  p.taps_h = (int)taps_hw[0]; p.taps_w = (int)taps_hw[1];
  p.box_n = (int)box_nhw[0]; p.box_h = (int)box_nhw[1]; p.box_w = (int)box_nhw[2];
  p.fp16 = dy.scalar_type() == at::kHalf ? 1 : 0;
  TORCH_CHECK(dw.numel() == (int64_t)p.m_total * p.taps_h * p.taps_w * p.c_total, "dw has the wrong size");
  p.dw = dw.data_ptr<float>();
  p.split_k = (int)split_k; p.accumulate = accumulate ? 1 : 0;
  if (slab.has_value()) {
    check_cuda(*slab, "slab", at::kFloat);
    TORCH_CHECK(multitap && slab->is_contiguous() && slab->numel() >= std::max<int64_t>(split_k, 1) * dw.numel(), "slab too small");
    p.slab = slab->data_ptr<float>();
  }
  if (multitap) ddpt::launch_wgrad_gemm_multitap(p, (int)block_n, current_stream());
  else ddpt::launch_wgrad_gemm(p, (int)block_n, current_stream());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "ddpt_b200 sm_100a kernels";
  m.def("multi_tensor_scale", &multi_tensor_scale);
  m.def("multi_tensor_axpby", &multi_tensor_axpby);
  m.def("unscale_arena", &unscale_arena);
  m.def("update_scale", &update_scale);
  m.def("fused_lookahead_sgd", &fused_lookahead_sgd);
  m.def("cast_bf16", &cast_bf16);
  m.def("conv_gemm", &conv_gemm, py::arg("a"), py::arg("b"), py::arg("out"), py::arg("out_nhw"), py::arg("box_nhw"),
        py::arg("taps_hw"), py::arg("tap_sign"), py::arg("a_off_hw"), py::arg("b_mn_major"), py::arg("bias"),
        py::arg("relu"), py::arg("stat_partials"), py::arg("split_k"), py::arg("block_n"), py::arg("mode"),
        py::arg("split_slabs") = false);
  m.def("splitk_finish", [](at::Tensor slabs, int64_t splits, at::Tensor out, c10::optional<at::Tensor> stat_partials) {
    TORCH_CHECK(slabs.is_cuda() && slabs.scalar_type() == at::kFloat && slabs.is_contiguous() && slabs.dim() == 2);
    TORCH_CHECK(out.is_cuda() && (out.scalar_type() == at::kBFloat16 || out.scalar_type() == at::kHalf) && out.is_contiguous() && out.dim() == 2);
    const int64_t rows = out.size(0), cols = out.size(1);
    TORCH_CHECK(slabs.size(0) == splits * rows && slabs.size(1) == cols, "slabs must be [splits * rows, cols]");
    float* part = nullptr;
    if (stat_partials.has_value()) {
      TORCH_CHECK(stat_partials->scalar_type() == at::kFloat && stat_partials->numel() >= (int64_t)ddpt::splitk_finish_max_rows() * 2 * cols);
      part = stat_partials->data_ptr<float>();
    }
    c10::cuda::CUDAGuard guard(out.device());
    return (int64_t)ddpt::launch_splitk_finish(slabs.data_ptr<float>(), (int)splits, rows * cols, (int)rows, (int)cols,
                                               reinterpret_cast<__nv_bfloat16*>(out.data_ptr()), part,
                                               out.scalar_type() == at::kHalf ? 1 : 0, current_stream());
  });
  m.def("wgrad_gemm", &wgrad_gemm);
  ddpt_bind_extra(m);
}
