"""CPU checks of the two algebraic rewrites the round-2 kernels rely on (the CUDA implementations are tested on the GPU in
test_gpu_kernels.py / test_gpu_toynet.py; these pin down the MATH in plain PyTorch so it can be checked anywhere).

1. conv_dgrad_kwmerge.cu: the data gradient of a 3x3 convolution computed with the three filter COLUMNS merged into
   the GEMM's N dimension and summed with a pixel shift afterwards.
2. toynet_fused.py::_folds_bias: a convolution bias in front of a training-mode BatchNorm can be left out of the
   convolution if running_mean (training) and the eval-mode shift are corrected."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


@pytest.mark.parametrize("n,h,w,co,pad", [(2, 26, 26, 16, 0), (1, 9, 37, 8, 0), (2, 14, 14, 8, 1)])
def test_merged_column_data_gradient(n, h, w, co, pad):
    """dx[y, x] = sum_kw D_kw[y, x - kw]  with  D_kw[y, c] = sum_kh dy[y - kh + pad, c + pad - ... ] W[kh, kw]:
    per tile of 4 rows x 32 lanes exactly as the kernel walks it (lane L <-> dy column x0 + pad - 2 + L, outputs on
    lanes 2..31, out-of-image dy reads are zero)."""
    ci = 4
    g = torch.Generator().manual_seed(0)
    ho, wo = h - 2 + 2 * pad, w - 2 + 2 * pad
    dy = torch.randn(n, ho, wo, co, generator=g, dtype=torch.float64)
    wk = torch.randn(co, 3, 3, ci, generator=g, dtype=torch.float64)                  # [co][kh][kw][ci], as the kernels store it
    want = F.conv_transpose2d(dy.permute(0, 3, 1, 2), wk.permute(0, 3, 1, 2), padding=pad).permute(0, 2, 3, 1)

    def dy_at(img, y, x):            # TMA zero fill outside the image
        if 0 <= y < ho and 0 <= x < wo:
            return dy[img, y, x]
        return torch.zeros(co, dtype=torch.float64)

    out = torch.full((n, h, w, ci), float("nan"), dtype=torch.float64)
    for img in range(n):
        for y0 in range(0, h, 4):
            for x0 in range(0, w, 30):
                # accumulator: 4 rows x 32 lanes x (kw, ci); the three kh taps accumulate through row-shifted views
                acc = torch.zeros(4, 32, 3, ci, dtype=torch.float64)
                for r in range(4):
                    for lane in range(32):
                        col = x0 + pad - 2 + lane
                        for kh in range(3):
                            v = dy_at(img, y0 + r + pad - kh, col)              # brick row r + dr, dr = 2 - kh
                            acc[r, lane] += torch.einsum("o,oki->ki", v, wk[:, kh])
                for r in range(4):
                    for lane in range(2, 32):
                        y, x = y0 + r, x0 + lane - 2
                        if y < h and x < w:
                            out[img, y, x] = acc[r, lane, 0] + acc[r, lane - 1, 1] + acc[r, lane - 2, 2]
    assert not torch.isnan(out).any()                                               # every output written exactly once
    torch.testing.assert_close(out, want, rtol=1e-10, atol=1e-10)


def test_conv_bias_folds_into_batchnorm_statistics():
    torch.manual_seed(0)
    conv, bn = nn.Conv2d(3, 8, 3, bias=True).double(), nn.BatchNorm2d(8).double()
    with torch.no_grad():
        conv.bias.uniform_(-1, 1); bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    x = torch.randn(16, 3, 10, 10, dtype=torch.float64, requires_grad=True)
    # reference: training step, then eval on new data
    bn.train()
    z = torch.relu(bn(conv(x)))
    z.square().sum().backward()
    ref = dict(z=z.detach(), dx=x.grad.clone(), dw=conv.weight.grad.clone(), db=conv.bias.grad.clone(),
               dgamma=bn.weight.grad.clone(), dbeta=bn.bias.grad.clone(), rm=bn.running_mean.clone(),
               rv=bn.running_var.clone())
    bn.eval()
    x2 = torch.randn(4, 3, 10, 10, dtype=torch.float64)
    ref_eval = torch.relu(bn(conv(x2))).detach()

    # folded: the convolution never adds its bias; the statistics carry it where it is observable
    b = conv.bias.detach()
    xf = x.detach().clone().requires_grad_(True)
    w = conv.weight.detach().clone().requires_grad_(True)
    gamma, beta = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    y = F.conv2d(xf, w)                                                            # bias-free activations (what is stored)
    mean, var = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=False)
    scale = gamma / torch.sqrt(var + bn.eps)
    zf = torch.relu(y * scale[None, :, None, None] + (beta - mean * scale)[None, :, None, None])
    zf.square().sum().backward()
    count = y.numel() / y.shape[1]
    rm = (1 - bn.momentum) * torch.zeros(8, dtype=torch.float64) + bn.momentum * (mean.detach() + b)      # bn_finalize(conv_bias=b)
    rv = (1 - bn.momentum) * torch.ones(8, dtype=torch.float64) + bn.momentum * var.detach() * count / (count - 1)
    torch.testing.assert_close(zf.detach(), ref["z"]); torch.testing.assert_close(xf.grad, ref["dx"])
    torch.testing.assert_close(w.grad, ref["dw"]); torch.testing.assert_close(gamma.grad, ref["dgamma"])
    torch.testing.assert_close(beta.grad, ref["dbeta"])
    torch.testing.assert_close(rm, ref["rm"]); torch.testing.assert_close(rv, ref["rv"])
    assert float(ref["db"].abs().max()) < 1e-9                  # the bias gradient is exactly zero: nothing to compute
    # eval: shift = beta - (running_mean - b) * scale   (bn_eval_coeffs(conv_bias=b))
    sc = gamma.detach() / torch.sqrt(rv + bn.eps)
    sh = beta.detach() - (rm - b) * sc
    got_eval = torch.relu(F.conv2d(x2, w.detach()) * sc[None, :, None, None] + sh[None, :, None, None])
    torch.testing.assert_close(got_eval, ref_eval)
