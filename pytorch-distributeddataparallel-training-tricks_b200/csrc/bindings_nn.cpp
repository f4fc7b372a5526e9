// Bindings for the bandwidth-bound Toy_Net kernels (csrc/nn/toynet_elementwise.cu).
#include <cstdlib>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cuda_fp16.h>
#include "include/ddpt_toynet.h"
#include "include/ddpt_fp8.h"

namespace {

cudaStream_t cur() { return at::cuda::getCurrentCUDAStream().stream(); }
const __nv_bfloat16* bf(const at::Tensor& t) { return reinterpret_cast<const __nv_bfloat16*>(t.data_ptr()); }
__nv_bfloat16* bfm(at::Tensor& t) { return reinterpret_cast<__nv_bfloat16*>(t.data_ptr()); }

void need(const at::Tensor& t, at::ScalarType dt, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == dt && t.is_contiguous(), name, ": expected contiguous CUDA ", dt);
}
// activations are bf16 (O1) or fp16 (O1-fp16); every activation tensor of one call must have the same type
bool is_half(const at::Tensor& t) { return t.scalar_type() == at::kHalf; }
void need_act(const at::Tensor& t, const at::Tensor& like, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && (t.scalar_type() == at::kBFloat16 || t.scalar_type() == at::kHalf) &&
              t.scalar_type() == like.scalar_type(), name, ": expected a contiguous CUDA bf16 / fp16 tensor (all of one type)");
}
const __half* hf(const at::Tensor& t) { return reinterpret_cast<const __half*>(t.data_ptr()); }
__half* hfm(at::Tensor& t) { return reinterpret_cast<__half*>(t.data_ptr()); }

int64_t conv0_partial_rows() { return ddpt::conv0_partial_rows(); }
int64_t bn_bwd_partial_rows() { return ddpt::bn_bwd_partial_rows(); }

int64_t conv0_fwd(at::Tensor x, at::Tensor weight, at::Tensor bias, at::Tensor y, c10::optional<at::Tensor> partials) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && (x.scalar_type() == at::kByte || x.scalar_type() == at::kFloat));
  TORCH_CHECK(x.numel() % (28 * 28) == 0, "conv0 expects 28x28 single-channel images");
  need(weight, at::kFloat, "weight"); need(bias, at::kFloat, "bias"); need_act(y, y, "y");
  const int batch = (int)(x.numel() / (28 * 28));
  TORCH_CHECK(weight.numel() == 64 * 9 && bias.numel() == 64 && y.numel() == (int64_t)batch * 26 * 26 * 64);
  c10::cuda::CUDAGuard g(x.device());
  float* part = nullptr;
  double* accum = nullptr;
  if (partials.has_value() && partials->scalar_type() == at::kDouble) {     // fp64 accumulators [2][64]
    need(*partials, at::kDouble, "partials");
    TORCH_CHECK(partials->numel() >= 2 * 64);
    accum = partials->data_ptr<double>();
  } else if (partials.has_value()) {
    need(*partials, at::kFloat, "partials");
    TORCH_CHECK(partials->numel() >= ddpt::conv0_partial_rows() * 2 * 64);
    part = partials->data_ptr<float>();
  }
  if (is_half(y))
    return ddpt::launch_conv0_fwd(x.data_ptr(), x.scalar_type() == at::kByte, weight.data_ptr<float>(),
                                  bias.data_ptr<float>(), hfm(y), part, accum, batch, cur());
  return ddpt::launch_conv0_fwd(x.data_ptr(), x.scalar_type() == at::kByte, weight.data_ptr<float>(),
                                bias.data_ptr<float>(), bfm(y), part, accum, batch, cur());
}

void conv0_wgrad(at::Tensor x, at::Tensor dy, at::Tensor dw, at::Tensor db, c10::optional<at::Tensor> partials,
                 c10::optional<at::Tensor> bn_y, c10::optional<at::Tensor> bn_g, c10::optional<at::Tensor> bn_coeffs,
                 c10::optional<at::Tensor> bn_bwd_coef, c10::optional<bool> tensor_core, int64_t channels_per_thread) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && (x.scalar_type() == at::kByte || x.scalar_type() == at::kFloat));
  need_act(dy, dy, "dy"); need(dw, at::kFloat, "dw"); need(db, at::kFloat, "db");
  const int batch = (int)(x.numel() / (28 * 28));
  TORCH_CHECK(dy.numel() == (int64_t)batch * 26 * 26 * 64 && dw.numel() == 576 && db.numel() == 64);
  c10::cuda::CUDAGuard g(x.device());
  float* part = nullptr;
  if (partials.has_value()) {
    need(*partials, at::kFloat, "partials");
    TORCH_CHECK(partials->numel() >= (int64_t)ddpt::conv0_wgrad_partial_rows() * 640, "conv0_wgrad: partials too small");
    part = partials->data_ptr<float>();
  }
  ddpt::Conv0BnFuse bn{};
  if (bn_y.has_value()) {         // fused BatchNorm-backward apply: dy is derived from (y0, g) on the fly
    TORCH_CHECK(bn_g.has_value() && bn_coeffs.has_value() && bn_bwd_coef.has_value());
    need_act(*bn_y, dy, "bn_y"); need_act(*bn_g, dy, "bn_g");
    need(*bn_coeffs, at::kFloat, "bn_coeffs"); need(*bn_bwd_coef, at::kFloat, "bn_bwd_coef");
    TORCH_CHECK(bn_y->numel() == dy.numel() && bn_g->numel() == dy.numel() && bn_coeffs->numel() >= 4 * 64 && bn_bwd_coef->numel() >= 3 * 64);
    const float* co = bn_coeffs->data_ptr<float>();
    bn = ddpt::Conv0BnFuse{bn_y->data_ptr(), bn_g->data_ptr(), co, co + 64, co + 128, co + 192, bn_bwd_coef->data_ptr<float>()};
  }
  // tensor-core variant (conv0_wgrad_tc.cu): opt-in (argument, or DDPT_CONV0_WGRAD_TC=1).  Measured inside the graphed step
  // it is no faster than the CUDA-core kernel (129 vs 114 us, profiles/step_timeline_r2.txt): producing the swizzled
  // operands costs the threads as many instructions as the 80 FFMA per gradient vector they replace.
  static const bool env_tc = std::getenv("DDPT_CONV0_WGRAD_TC") != nullptr && std::atoi(std::getenv("DDPT_CONV0_WGRAD_TC")) != 0;
  if (tensor_core.has_value() ? *tensor_core : env_tc) {
    ddpt::launch_conv0_wgrad_tc(x.data_ptr(), x.scalar_type() == at::kByte, dy.data_ptr(), is_half(dy) ? 1 : 0,
                                dw.data_ptr<float>(), db.data_ptr<float>(), part, batch, bn, cur());
    return;
  }
  if (is_half(dy))
    ddpt::launch_conv0_wgrad(x.data_ptr(), x.scalar_type() == at::kByte, hf(dy), dw.data_ptr<float>(),
                             db.data_ptr<float>(), part, batch, bn, cur(), (int)channels_per_thread);
  else
    ddpt::launch_conv0_wgrad(x.data_ptr(), x.scalar_type() == at::kByte, bf(dy), dw.data_ptr<float>(),
                             db.data_ptr<float>(), part, batch, bn, cur(), (int)channels_per_thread);
}

// coeffs: float [4][C] = scale, shift, mean, invstd
void bn_finalize(at::Tensor partials, int64_t num_partials, double count, at::Tensor gamma, at::Tensor beta,
                 double eps, double momentum, at::Tensor coeffs, c10::optional<at::Tensor> running_mean,
                 c10::optional<at::Tensor> running_var, c10::optional<at::Tensor> nbt,
                 c10::optional<at::Tensor> conv_bias) {
  need(partials, at::kFloat, "partials"); need(gamma, at::kFloat, "gamma"); need(beta, at::kFloat, "beta");
  need(coeffs, at::kFloat, "coeffs");
  const int c = (int)gamma.numel();
  TORCH_CHECK(coeffs.numel() >= 4 * c && partials.numel() >= num_partials * 2 * c);
  c10::cuda::CUDAGuard g(gamma.device());
  float* co = coeffs.data_ptr<float>();
  ddpt::BnCoeffs out{co, co + c, co + 2 * c, co + 3 * c};
  const float* cb = nullptr;
  if (conv_bias.has_value()) { need(*conv_bias, at::kFloat, "conv_bias"); TORCH_CHECK(conv_bias->numel() == c); cb = conv_bias->data_ptr<float>(); }
  long long* n = nullptr;
  if (nbt.has_value()) { TORCH_CHECK(nbt->scalar_type() == at::kLong); n = reinterpret_cast<long long*>(nbt->data_ptr<int64_t>()); }
  ddpt::launch_bn_finalize(partials.data_ptr<float>(), (int)num_partials, c, count, gamma.data_ptr<float>(),
                           beta.data_ptr<float>(), (float)eps, (float)momentum, out,
                           running_mean.has_value() ? running_mean->data_ptr<float>() : nullptr,
                           running_var.has_value() ? running_var->data_ptr<float>() : nullptr, n, cb, cur());
}

void bn_eval_coeffs(at::Tensor gamma, at::Tensor beta, at::Tensor running_mean, at::Tensor running_var, double eps,
                    at::Tensor coeffs, c10::optional<at::Tensor> conv_bias) {
  need(gamma, at::kFloat, "gamma"); need(beta, at::kFloat, "beta"); need(running_mean, at::kFloat, "running_mean");
  need(running_var, at::kFloat, "running_var"); need(coeffs, at::kFloat, "coeffs");
  const int c = (int)gamma.numel();
  TORCH_CHECK(coeffs.numel() >= 4 * c);
  c10::cuda::CUDAGuard g(gamma.device());
  float* co = coeffs.data_ptr<float>();
  ddpt::launch_bn_eval_coeffs(gamma.data_ptr<float>(), beta.data_ptr<float>(), running_mean.data_ptr<float>(),
                              running_var.data_ptr<float>(), (float)eps, c, co, co + c,
                              conv_bias.has_value() ? conv_bias->data_ptr<float>() : nullptr, cur());
}

// accum (optional, fp64 [2][C]): fused statistics -- the kernel derives scale/shift itself, writes them (+ mean, invstd)
// into `coeffs` for the backward pass and updates running_mean / running_var / num_batches_tracked.
void bn_relu_pool(at::Tensor y, at::Tensor coeffs, at::Tensor out, bool pool, bool flatten_chw,
                  c10::optional<at::Tensor> q_out, c10::optional<at::Tensor> q_scale,
                  c10::optional<at::Tensor> accum, double count, c10::optional<at::Tensor> gamma,
                  c10::optional<at::Tensor> beta, double eps, double momentum,
                  c10::optional<at::Tensor> running_mean, c10::optional<at::Tensor> running_var,
                  c10::optional<at::Tensor> nbt) {
  need_act(y, y, "y"); need(coeffs, at::kFloat, "coeffs"); need_act(out, y, "out");
  TORCH_CHECK(y.dim() == 4);
  const int n = (int)y.size(0), h = (int)y.size(1), w = (int)y.size(2), c = (int)y.size(3);
  TORCH_CHECK(out.numel() == (pool ? y.numel() / 4 : y.numel()));
  c10::cuda::CUDAGuard g(y.device());
  float* co = coeffs.data_ptr<float>();
  uint8_t* qd = nullptr; uint8_t* qs = nullptr;
  if (q_out.has_value()) {
    TORCH_CHECK(q_scale.has_value());
    need(*q_out, at::kByte, "q_out"); need(*q_scale, at::kByte, "q_scale");
    TORCH_CHECK(q_out->numel() == out.numel() && q_scale->numel() == out.numel() / 32);
    qd = q_out->data_ptr<uint8_t>(); qs = q_scale->data_ptr<uint8_t>();
  }
  ddpt::BnFusedFwd f{};
  if (accum.has_value()) {
    need(*accum, at::kDouble, "accum"); TORCH_CHECK(accum->numel() >= 2 * c && gamma.has_value() && beta.has_value());
    need(*gamma, at::kFloat, "gamma"); need(*beta, at::kFloat, "beta");
    TORCH_CHECK(gamma->numel() == c && beta->numel() == c && coeffs.numel() >= 4 * c && count > 0);
    f.accum = accum->data_ptr<double>(); f.count = count; f.inv_count = 1.0 / count;
    f.unbias = count > 1.0 ? (float)(count / (count - 1.0)) : 1.0f;
    f.gamma = gamma->data_ptr<float>(); f.beta = beta->data_ptr<float>();
    f.eps = (float)eps; f.momentum = (float)momentum;
    f.out = ddpt::BnCoeffs{co, co + c, co + 2 * c, co + 3 * c};
    if (running_mean.has_value()) {
      TORCH_CHECK(running_var.has_value());
      need(*running_mean, at::kFloat, "running_mean"); need(*running_var, at::kFloat, "running_var");
      f.running_mean = running_mean->data_ptr<float>(); f.running_var = running_var->data_ptr<float>();
    }
    if (nbt.has_value()) { TORCH_CHECK(nbt->scalar_type() == at::kLong); f.num_batches_tracked = reinterpret_cast<long long*>(nbt->data_ptr<int64_t>()); }
  }
  if (is_half(y)) ddpt::launch_bn_relu_pool(hf(y), co, co + c, hfm(out), n, h, w, c, pool, flatten_chw, qd, qs, f, cur());
  else ddpt::launch_bn_relu_pool(bf(y), co, co + c, bfm(out), n, h, w, c, pool, flatten_chw, qd, qs, f, cur());
}

void chw_to_hwc(at::Tensor in, at::Tensor out, int64_t c, int64_t p) {
  need_act(in, in, "in"); need_act(out, in, "out");
  TORCH_CHECK(in.numel() == out.numel() && in.numel() % (c * p) == 0, "chw_to_hwc: [N, C*P] -> [N, P, C]");
  c10::cuda::CUDAGuard g(in.device());
  if (is_half(in)) ddpt::launch_chw_to_hwc(hf(in), hfm(out), (int)(in.numel() / (c * p)), (int)c, (int)p, cur());
  else ddpt::launch_chw_to_hwc(bf(in), bfm(out), (int)(in.numel() / (c * p)), (int)c, (int)p, cur());
}

int64_t bn_bwd_reduce(at::Tensor y, at::Tensor g_in, at::Tensor coeffs, at::Tensor partials, bool pool, bool flatten_chw) {
  need_act(y, y, "y"); need_act(g_in, y, "g"); need(coeffs, at::kFloat, "coeffs");
  const int n = (int)y.size(0), h = (int)y.size(1), w = (int)y.size(2), c = (int)y.size(3);
  TORCH_CHECK(g_in.numel() == (pool ? y.numel() / 4 : y.numel()));
  float* part = nullptr; double* accum = nullptr;
  if (partials.scalar_type() == at::kDouble) {            // fp64 accumulators [2][C] (zeroed by the caller)
    need(partials, at::kDouble, "partials"); TORCH_CHECK(partials.numel() >= 2 * c);
    accum = partials.data_ptr<double>();
  } else {
    need(partials, at::kFloat, "partials");
    TORCH_CHECK(partials.numel() >= (int64_t)ddpt::bn_bwd_partial_rows() * 2 * c);
    part = partials.data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(y.device());
  const float* co = coeffs.data_ptr<float>();
  if (is_half(y))
    return ddpt::launch_bn_bwd_reduce(hf(y), hf(g_in), co, co + c, co + 2 * c, co + 3 * c, part, accum, n,
                                      h, w, c, pool, flatten_chw, cur());
  return ddpt::launch_bn_bwd_reduce(bf(y), bf(g_in), co, co + c, co + 2 * c, co + 3 * c, part, accum, n,
                                    h, w, c, pool, flatten_chw, cur());
}

void bn_bwd_finalize(at::Tensor partials, int64_t num_partials, double count, at::Tensor gamma, at::Tensor coeffs,
                     at::Tensor dgamma, at::Tensor dbeta, at::Tensor bwd_coef, bool accumulate) {
  need(partials, at::kFloat, "partials"); need(gamma, at::kFloat, "gamma"); need(coeffs, at::kFloat, "coeffs");
  need(dgamma, at::kFloat, "dgamma"); need(dbeta, at::kFloat, "dbeta"); need(bwd_coef, at::kFloat, "bwd_coef");
  const int c = (int)gamma.numel();
  TORCH_CHECK(bwd_coef.numel() >= 3 * c && dgamma.numel() == c && dbeta.numel() == c);
  c10::cuda::CUDAGuard guard(gamma.device());
  ddpt::launch_bn_bwd_finalize(partials.data_ptr<float>(), (int)num_partials, c, count, gamma.data_ptr<float>(),
                               coeffs.data_ptr<float>() + 2 * c, coeffs.data_ptr<float>() + 3 * c, dgamma.data_ptr<float>(),
                               dbeta.data_ptr<float>(),
                               bwd_coef.data_ptr<float>(), accumulate, cur());
}

// accum (optional, fp64 [2][C] from bn_bwd_reduce): fused statistics -- the kernel derives its coefficients itself and
// writes dgamma / dbeta (bwd_coef is then unused).
void bn_bwd_apply(at::Tensor y, at::Tensor g_in, at::Tensor coeffs, at::Tensor bwd_coef, at::Tensor dy,
                  c10::optional<at::Tensor> dbias, bool pool, bool flatten_chw, c10::optional<at::Tensor> accum,
                  double count, c10::optional<at::Tensor> gamma, c10::optional<at::Tensor> dgamma,
                  c10::optional<at::Tensor> dbeta) {
  need_act(y, y, "y"); need_act(g_in, y, "g"); need(coeffs, at::kFloat, "coeffs");
  need(bwd_coef, at::kFloat, "bwd_coef"); need_act(dy, y, "dy");
  const int n = (int)y.size(0), h = (int)y.size(1), w = (int)y.size(2), c = (int)y.size(3);
  TORCH_CHECK(dy.numel() == y.numel());
  c10::cuda::CUDAGuard guard(y.device());
  const float* co = coeffs.data_ptr<float>();
  float* db = nullptr;
  if (dbias.has_value()) { need(*dbias, at::kFloat, "dbias"); TORCH_CHECK(dbias->numel() == c); db = dbias->data_ptr<float>(); }
  ddpt::BnFusedBwd f{};
  if (accum.has_value()) {
    need(*accum, at::kDouble, "accum");
    TORCH_CHECK(accum->numel() >= 2 * c && gamma.has_value() && dgamma.has_value() && dbeta.has_value() && count > 0);
    need(*gamma, at::kFloat, "gamma"); need(*dgamma, at::kFloat, "dgamma"); need(*dbeta, at::kFloat, "dbeta");
    TORCH_CHECK(gamma->numel() == c && dgamma->numel() == c && dbeta->numel() == c);
    f.accum = accum->data_ptr<double>(); f.count = count; f.inv_count = 1.0 / count; f.gamma = gamma->data_ptr<float>();
    f.dgamma = dgamma->data_ptr<float>(); f.dbeta = dbeta->data_ptr<float>();
  }
  if (is_half(y))
    ddpt::launch_bn_bwd_apply(hf(y), hf(g_in), co, co + c, co + 2 * c, co + 3 * c, bwd_coef.data_ptr<float>(), hfm(dy), db,
                              f, n, h, w, c, pool, flatten_chw, cur());
  else
    ddpt::launch_bn_bwd_apply(bf(y), bf(g_in), co, co + c, co + 2 * c, co + 3 * c, bwd_coef.data_ptr<float>(), bfm(dy), db,
                              f, n, h, w, c, pool, flatten_chw, cur());
}

void transpose_taps(at::Tensor w, at::Tensor wt, int64_t taps) {
  need_act(w, w, "w"); need_act(wt, w, "wt");
  TORCH_CHECK(w.dim() == 2 && wt.dim() == 2 && w.numel() == wt.numel() && w.size(1) % taps == 0 && wt.size(1) == taps * w.size(0),
              "transpose_taps: w [c_out, taps*c_in] -> wt [c_in, taps*c_out]");
  c10::cuda::CUDAGuard g(w.device());
  if (is_half(w)) ddpt::launch_transpose_taps(hf(w), hfm(wt), (int)w.size(0), (int)taps, (int)(w.size(1) / taps), cur());
  else ddpt::launch_transpose_taps(bf(w), bfm(wt), (int)w.size(0), (int)taps, (int)(w.size(1) / taps), cur());
}

// Forward-only when da is None.  dlogits: if given, used instead of the fused cross-entropy gradient.
void head(at::Tensor a, at::Tensor w, at::Tensor b, c10::optional<at::Tensor> target, c10::optional<at::Tensor> logits,
          c10::optional<at::Tensor> metrics, c10::optional<at::Tensor> loss_scale, c10::optional<at::Tensor> da,
          c10::optional<at::Tensor> dw, c10::optional<at::Tensor> db, c10::optional<at::Tensor> dlogits,
          double extra_div, c10::optional<at::Tensor> partials) {
  need_act(a, a, "a"); need(w, at::kFloat, "w"); need(b, at::kFloat, "b");
  TORCH_CHECK(a.dim() == 2 && a.size(1) == 512 && w.numel() == 5120 && b.numel() == 10);
  c10::cuda::CUDAGuard guard(a.device());
  ddpt::HeadArgs h{};
  h.a = a.data_ptr(); h.w = w.data_ptr<float>(); h.b = b.data_ptr<float>(); h.batch = (int)a.size(0);
  if (target.has_value()) { TORCH_CHECK(target->scalar_type() == at::kLong && target->numel() == h.batch);
    h.target = reinterpret_cast<const long long*>(target->data_ptr<int64_t>()); }
  if (logits.has_value()) { need(*logits, at::kFloat, "logits"); h.logits = logits->data_ptr<float>(); }
  if (metrics.has_value()) { need(*metrics, at::kFloat, "metrics"); h.metrics = metrics->data_ptr<float>(); }
  if (loss_scale.has_value()) h.loss_scale = loss_scale->data_ptr<float>();
  if (da.has_value()) {
    TORCH_CHECK(dw.has_value() && db.has_value());
    need_act(*da, a, "da"); need(*dw, at::kFloat, "dw"); need(*db, at::kFloat, "db");
    h.da = da->data_ptr(); h.dw = dw->data_ptr<float>(); h.db = db->data_ptr<float>();
  }
  if (dlogits.has_value()) { need(*dlogits, at::kFloat, "dlogits"); h.dlogits = dlogits->data_ptr<float>(); }
  TORCH_CHECK(h.target != nullptr || (h.metrics == nullptr && (h.da == nullptr || h.dlogits != nullptr)),
              "head: target is required for metrics / fused cross-entropy gradient");
  h.extra_div = (float)extra_div;
  if (partials.has_value()) {
    need(*partials, at::kFloat, "partials");
    TORCH_CHECK(partials->numel() >= (int64_t)at::cuda::getCurrentDeviceProperties()->multiProcessorCount * ddpt::head_partial_cols(),
                "head: partials too small");
    h.partials = partials->data_ptr<float>();
  }
  if (is_half(a)) ddpt::launch_head_f16(h, cur());
  else ddpt::launch_head(h, cur());
}

// x: [rows, cols] bf16/fp32 contiguous (cols % 32 == 0) -> (uint8 e4m3 [rows, cols], uint8 UE8M0 [rows, cols/32])
std::pair<at::Tensor, at::Tensor> quantize_mxfp8(at::Tensor x, c10::optional<at::Tensor> data_out,
                                                 c10::optional<at::Tensor> scale_out) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && (x.scalar_type() == at::kBFloat16 || x.scalar_type() == at::kFloat));
  const int64_t cols = x.size(-1), rows = x.numel() / cols;
  TORCH_CHECK(cols % 32 == 0, "quantize_mxfp8: the last dimension must be a multiple of 32");
  c10::cuda::CUDAGuard g(x.device());
  auto shape = x.sizes().vec();
  at::Tensor data = data_out.has_value() ? *data_out : at::empty(shape, x.options().dtype(at::kByte));
  shape.back() = cols / 32;
  at::Tensor scale = scale_out.has_value() ? *scale_out : at::empty(shape, x.options().dtype(at::kByte));
  TORCH_CHECK(data.scalar_type() == at::kByte && data.numel() == x.numel() && data.is_contiguous());
  TORCH_CHECK(scale.scalar_type() == at::kByte && scale.numel() == x.numel() / 32 && scale.is_contiguous());
  ddpt::launch_quantize_mxfp8(x.data_ptr(), x.scalar_type() == at::kFloat, data.data_ptr<uint8_t>(),
                              scale.data_ptr<uint8_t>(), rows, (int)cols, cur());
  return {data, scale};
}

int64_t conv_gemm_mxfp8(at::Tensor a, at::Tensor a_scale, at::Tensor b, at::Tensor b_scale, at::Tensor out,
                        std::vector<int64_t> out_nhw, std::vector<int64_t> box_nhw, std::vector<int64_t> taps_hw,
                        c10::optional<at::Tensor> bias, bool relu, c10::optional<at::Tensor> stat_partials) {
  need(a, at::kByte, "a"); need(a_scale, at::kByte, "a_scale"); need(b, at::kByte, "b"); need(b_scale, at::kByte, "b_scale");
  need(out, at::kBFloat16, "out");
  TORCH_CHECK(a.dim() == 4 && b.dim() == 2 && out.dim() == 2);
  c10::cuda::CUDAGuard g(a.device());
  ddpt::Mxfp8GemmProblem p{};
  p.a = a.data_ptr<uint8_t>(); p.a_scale = a_scale.data_ptr<uint8_t>();
  p.a_n = (int)a.size(0); p.a_h = (int)a.size(1); p.a_w = (int)a.size(2); p.a_c = (int)a.size(3);
  p.out_n = (int)out_nhw[0]; p.out_h = (int)out_nhw[1]; p.out_w = (int)out_nhw[2];
  p.box_n = (int)box_nhw[0]; p.box_h = (int)box_nhw[1]; p.box_w = (int)box_nhw[2];
  p.taps_h = (int)taps_hw[0]; p.taps_w = (int)taps_hw[1];
  const int64_t taps = p.taps_h * p.taps_w;
  TORCH_CHECK(a_scale.numel() == a.numel() / 32 && b.size(1) == taps * p.a_c && b_scale.numel() == b.numel() / 32);
  p.b = b.data_ptr<uint8_t>(); p.b_scale = b_scale.data_ptr<uint8_t>(); p.n_total = (int)b.size(0);
  TORCH_CHECK(out.size(0) >= (int64_t)p.out_n * p.out_h * p.out_w && out.size(1) >= p.n_total);
  p.out = out.data_ptr(); p.ld_out = out.stride(0);
  if (bias.has_value()) { need(*bias, at::kFloat, "bias"); p.bias = bias->data_ptr<float>(); }
  p.relu = relu;
  if (stat_partials.has_value()) {
    need(*stat_partials, at::kFloat, "stat_partials");
    const int64_t tiles = (int64_t)((p.out_w + p.box_w - 1) / p.box_w) * ((p.out_h + p.box_h - 1) / p.box_h) *
                          ((p.out_n + p.box_n - 1) / p.box_n);
    const int64_t groups = std::max<int64_t>(tiles, std::min<int64_t>(tiles * (p.n_total / 128),
                                             at::cuda::getCurrentDeviceProperties()->multiProcessorCount));
    TORCH_CHECK(stat_partials->numel() >= groups * 8 * p.n_total, "stat_partials too small");
    p.stat_partials = stat_partials->data_ptr<float>();
  }
  return ddpt::launch_conv_gemm_mxfp8(p, cur());
}

}  // namespace

void ddpt_bind_nn(pybind11::module_& m) {
  m.def("quantize_mxfp8", &quantize_mxfp8);
  m.def("conv_gemm_mxfp8", &conv_gemm_mxfp8);
  m.def("conv0_partial_rows", &conv0_partial_rows);
  m.def("bn_bwd_partial_rows", &bn_bwd_partial_rows);
  m.def("conv0_fwd", &conv0_fwd);
  m.def("conv0_wgrad", &conv0_wgrad, py::arg("x"), py::arg("dy"), py::arg("dw"), py::arg("db"), py::arg("partials") = py::none(),
        py::arg("bn_y") = py::none(), py::arg("bn_g") = py::none(), py::arg("bn_coeffs") = py::none(),
        py::arg("bn_bwd_coef") = py::none(), py::arg("tensor_core") = py::none(), py::arg("channels_per_thread") = 0);
  m.def("transpose_taps", &transpose_taps);
  m.def("det_workspace_floats", []() {     // fp32 elements of the deterministic-mode partial-row workspace
    return (int64_t)std::max<int64_t>((int64_t)ddpt::conv0_wgrad_partial_rows() * 640,
                                      (int64_t)at::cuda::getCurrentDeviceProperties()->multiProcessorCount * ddpt::head_partial_cols());
  });
  m.def("bn_finalize", &bn_finalize, py::arg("partials"), py::arg("num_partials"), py::arg("count"), py::arg("gamma"),
        py::arg("beta"), py::arg("eps"), py::arg("momentum"), py::arg("coeffs"), py::arg("running_mean") = py::none(),
        py::arg("running_var") = py::none(), py::arg("nbt") = py::none(), py::arg("conv_bias") = py::none());
  m.def("bn_eval_coeffs", &bn_eval_coeffs, py::arg("gamma"), py::arg("beta"), py::arg("running_mean"), py::arg("running_var"),
        py::arg("eps"), py::arg("coeffs"), py::arg("conv_bias") = py::none());
  m.def("bn_relu_pool", &bn_relu_pool, py::arg("y"), py::arg("coeffs"), py::arg("out"), py::arg("pool"),
        py::arg("flatten_chw"), py::arg("q_out") = py::none(), py::arg("q_scale") = py::none(),
        py::arg("accum") = py::none(), py::arg("count") = 0.0, py::arg("gamma") = py::none(),
        py::arg("beta") = py::none(), py::arg("eps") = 1e-5, py::arg("momentum") = 0.1,
        py::arg("running_mean") = py::none(), py::arg("running_var") = py::none(), py::arg("nbt") = py::none());
  m.def("bn_bwd_reduce", &bn_bwd_reduce);
  m.def("bn_bwd_finalize", &bn_bwd_finalize);
  m.def("bn_bwd_apply", &bn_bwd_apply, py::arg("y"), py::arg("g"), py::arg("coeffs"), py::arg("bwd_coef"), py::arg("dy"),
        py::arg("dbias"), py::arg("pool"), py::arg("flatten_chw"), py::arg("accum") = py::none(), py::arg("count") = 0.0,
        py::arg("gamma") = py::none(), py::arg("dgamma") = py::none(), py::arg("dbeta") = py::none());
  m.def("chw_to_hwc", &chw_to_hwc);
  m.def("head", &head, py::arg("a"), py::arg("w"), py::arg("b"), py::arg("target"), py::arg("logits"), py::arg("metrics"),
        py::arg("loss_scale"), py::arg("da"), py::arg("dw"), py::arg("db"), py::arg("dlogits"), py::arg("extra_div"),
        py::arg("partials") = py::none());
}
