"""sm_100a kernel numerics vs plain PyTorch fp32 references (run on a B200: pytest -m gpu)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    from ddpt_b200 import ops
    return ops.require_native()


@pytest.fixture(params=["halo", "persistent", "pair", "per_tile"])
def gemm_flavour(request, monkeypatch):
    """Every GEMM/conv test runs against all forward/dgrad kernels (halo reuse, persistent, persistent as CTA pairs
    with cta_group::2 MMAs, one CTA per tile)."""
    monkeypatch.setenv("DDPT_GEMM_PERSISTENT", "0" if request.param == "per_tile" else "1")
    monkeypatch.setenv("DDPT_GEMM_HALO", "force" if request.param == "halo" else "0")
    monkeypatch.setenv("DDPT_GEMM_PAIR", "1" if request.param == "pair" else "0")
    monkeypatch.setenv("DDPT_DGRAD_KWMERGE", "0")      # the merged-column data gradient has its own tests below
    return request.param


def _bf16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _rel_err(got, want):
    return float((got.float() - want.float()).norm() / want.float().norm().clamp_min(1e-20))


# ---------------------------------------------------------------------------------------------
# tcgen05 GEMM / conv
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k,bn", [(128, 64, 64, 64), (256, 128, 256, 128), (1024, 512, 8192, 128),
                                      (304, 256, 512, 256), (1000, 64, 128, 64)])
def test_linear_fwd(ext, gemm_flavour, m, n, k, bn):
    from ddpt_b200.ops import gemm
    x, w = _bf16(m, k, seed=1), _bf16(n, k, scale=k ** -0.5, seed=2)
    bias = torch.randn(n, device="cuda")
    y, _ = gemm.linear_fwd(x, w, bias, block_n=bn, out_dtype=torch.float32)
    want = x.float() @ w.float().t() + bias
    assert _rel_err(y, want) < 2e-3
    yb, _ = gemm.linear_fwd(x, w, bias, relu=True, block_n=bn)
    assert _rel_err(yb, want.relu()) < 1e-2


def test_linear_fwd_split_k(ext, gemm_flavour):
    from ddpt_b200.ops import gemm
    x, w = _bf16(1024, 8192, seed=3), _bf16(512, 8192, scale=8192 ** -0.5, seed=4)
    y, _ = gemm.linear_fwd(x, w, None, out_dtype=torch.float32, split_k=4, block_n=128)
    assert _rel_err(y, x.float() @ w.float().t()) < 2e-3


@pytest.mark.parametrize("m,n,k,split,bn", [(1024, 512, 8192, 8, 256), (608, 512, 8192, 4, 128), (256, 256, 1024, 2, 256)])
def test_linear_fwd_splitk_slabs_and_statistics(ext, m, n, k, split, bn):
    """Split-K linear: fp32 slabs (no atomics) folded in split order, bf16 output, BatchNorm partial rows."""
    from ddpt_b200.ops import gemm
    x, w = _bf16(m, k, seed=51), _bf16(n, k, scale=k ** -0.5, seed=52)
    bias = torch.randn(n, device="cuda")
    part = torch.zeros(ext.bn_bwd_partial_rows() * 2 * n, device="cuda")
    y, rows = gemm.linear_fwd_splitk(x, w, bias, part, split_k=split, block_n=bn)
    want = x.float() @ w.float().t() + bias
    assert _rel_err(y, want) < 1e-2
    stats = part[: rows * 2 * n].view(rows, 2, n).double().sum(0)
    torch.testing.assert_close(stats[0], want.double().sum(0), rtol=2e-3, atol=0.05 * math.sqrt(m))
    torch.testing.assert_close(stats[1], (want.double() ** 2).sum(0), rtol=5e-3, atol=1e-2)
    y2, _ = gemm.linear_fwd_splitk(x, w, bias, None, split_k=split, block_n=bn)
    assert torch.equal(y, y2)                                    # ordered fold: bit-reproducible


@pytest.mark.parametrize("m,n,k", [(256, 128, 128), (1024, 8192, 512), (304, 512, 64)])
def test_linear_dgrad_mn_major_b(ext, gemm_flavour, m, n, k):
    """dx[M,K] = dy[M,N] @ w[N,K] -- B operand consumed MN-major."""
    from ddpt_b200.ops import gemm
    dy, w = _bf16(m, n, seed=5), _bf16(n, k, scale=n ** -0.5, seed=6)
    dx = gemm.linear_dgrad(dy, w)
    assert _rel_err(dx, dy.float() @ w.float()) < 1e-2


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (1024, 512, 8192), (304, 128, 256)])
def test_linear_wgrad_mn_major_both(ext, gemm_flavour, m, n, k):
    """dw[N,K] = dy[M,N]^T @ x[M,K] -- both operands MN-major, split-K reduction."""
    from ddpt_b200.ops import gemm
    dy, x = _bf16(m, n, seed=7), _bf16(m, k, seed=8)
    want = dy.float().t() @ x.float()
    for split in (1, None):
        for mt in (False, True):
            dw = gemm.linear_wgrad(dy, x, split_k=split, multitap=mt)
            assert _rel_err(dw, want) < 2e-3, (split, mt)


CONVS = [(8, 26, 26, 64, 128), (6, 12, 12, 128, 256), (4, 10, 10, 256, 512), (5, 7, 9, 64, 64)]


def _conv_ref(x_nhwc, w_km, c_out, c_in):
    w = w_km.float().view(c_out, 3, 3, c_in).permute(0, 3, 1, 2)
    return F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w)


@pytest.mark.parametrize("n,h,w,ci,co", CONVS)
def test_conv3x3_fwd_and_stats(ext, gemm_flavour, n, h, w, ci, co):
    from ddpt_b200.ops import gemm
    x, wk = _bf16(n, h, w, ci, seed=9), _bf16(co, 9 * ci, scale=(9 * ci) ** -0.5, seed=10)
    bias = torch.randn(co, device="cuda")
    want = _conv_ref(x, wk, co, ci) + bias.view(1, -1, 1, 1)
    y32, tiles = gemm.conv_fwd(x, wk, 3, 3, bias, out_dtype=torch.float32)
    assert _rel_err(y32.permute(0, 3, 1, 2), want) < 2e-3
    groups = gemm.stat_row_groups(n, h - 2, w - 2, co)
    assert groups >= tiles
    stats = torch.zeros(groups * 8 * co, device="cuda")
    y, _ = gemm.conv_fwd(x, wk, 3, 3, bias, stat_partials=stats)
    assert _rel_err(y.permute(0, 3, 1, 2), want) < 1e-2
    part = stats.view(groups * 4, 2, co).sum(0)
    flat = want.permute(0, 2, 3, 1).reshape(-1, co)
    torch.testing.assert_close(part[0], flat.sum(0), rtol=2e-3, atol=2e-2 * math.sqrt(flat.shape[0]))
    torch.testing.assert_close(part[1], (flat * flat).sum(0), rtol=2e-3, atol=1e-2)


@pytest.mark.parametrize("n,h,w,ci,co", CONVS)
def test_conv3x3_dgrad(ext, gemm_flavour, n, h, w, ci, co):
    from ddpt_b200.ops import gemm
    dy, wk = _bf16(n, h - 2, w - 2, co, seed=11), _bf16(co, 9 * ci, scale=(9 * co) ** -0.5, seed=12)
    wt = wk.float().view(co, 3, 3, ci).permute(0, 3, 1, 2)
    want = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt)
    dx = gemm.conv_dgrad(dy, wk, 3, 3)
    assert dx.shape == (n, h, w, ci)
    assert _rel_err(dx.permute(0, 3, 1, 2), want) < 1e-2


@pytest.mark.parametrize("n,h,w,ci,co", [(256, 10, 10, 256, 512), (300, 12, 12, 128, 256), (130, 9, 9, 64, 128),
                                         (200, 11, 7, 128, 128)])
def test_conv3x3_dgrad_skips_out_of_range_taps(ext, gemm_flavour, n, h, w, ci, co):
    """Batches >= 128 pick one-pixel bricks over 128 images; border pixels then run only the filter taps that reach
    them (persistent kernels, incl. the CTA-pair union / interleave logic and the resident-weight tap counters)."""
    from ddpt_b200.ops import gemm
    assert gemm.pick_box_dgrad(1024, 10, 10, 3, 3) == (128, 1, 1)
    dy, wk = _bf16(n, h - 2, w - 2, co, seed=21), _bf16(co, 9 * ci, scale=(9 * co) ** -0.5, seed=22)
    wt = wk.float().view(co, 3, 3, ci).permute(0, 3, 1, 2)
    want = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt)
    dx = gemm.conv_dgrad(dy, wk, 3, 3)
    assert _rel_err(dx.permute(0, 3, 1, 2), want) < 1e-2
    # per-pixel check: a skipped tap that was needed would zero out whole border rows/columns
    err = (dx.permute(0, 3, 1, 2).float() - want).abs().amax(dim=(0, 1))
    assert float(err.max()) < 0.05 * float(want.abs().max()), err


@pytest.mark.parametrize("n,h,w,co", [(8, 26, 26, 128), (130, 9, 11, 128), (256, 10, 10, 256), (5, 7, 9, 64)])
def test_conv3x3_dgrad_64_channels_as_cta_pair(ext, gemm_flavour, n, h, w, co):
    """C_in = 64: with the tap-transposed (K-major) weights the data gradient runs on cta_group::2 MMAs (256 x 64 x 16),
    each CTA staging 32 weight rows; same numbers as the MN-major single-CTA path."""
    from ddpt_b200.ops import gemm
    ci = 64
    dy, wk = _bf16(n, h - 2, w - 2, co, seed=41), _bf16(co, 9 * ci, scale=(9 * co) ** -0.5, seed=42)
    w_t = gemm.transpose_taps(wk, 9)
    assert torch.equal(w_t.view(ci, 9, co), wk.view(co, 9, ci).permute(2, 1, 0))
    wt = wk.float().view(co, 3, 3, ci).permute(0, 3, 1, 2)
    want = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt)
    os.environ["DDPT_DGRAD64_PAIR"] = "1"
    try:
        dx = gemm.conv_dgrad(dy, wk, 3, 3, w_t=w_t)
    finally:
        os.environ.pop("DDPT_DGRAD64_PAIR", None)
    assert _rel_err(dx.permute(0, 3, 1, 2), want) < 1e-2
    err = (dx.permute(0, 3, 1, 2).float() - want).abs().amax(dim=(0, 1))
    assert float(err.max()) < 0.05 * float(want.abs().max()), err


@pytest.mark.parametrize("n,h,w,ci,co", [(8, 26, 26, 64, 128), (6, 12, 12, 128, 256), (4, 10, 10, 256, 512),
                                         (3, 9, 13, 64, 128)])
def test_conv3x3_wgrad(ext, gemm_flavour, n, h, w, ci, co):
    from ddpt_b200.ops import gemm
    dy, x = _bf16(n, h - 2, w - 2, co, seed=13), _bf16(n, h, w, ci, seed=14)
    xg = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wref = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(xg, wref).backward(dy.float().permute(0, 3, 1, 2))
    want = wref.grad.permute(0, 2, 3, 1)                       # [co, kh, kw, ci]
    for split in (1, None):
        for mt in (False, True):
            dw = gemm.conv_wgrad(dy, x, 3, 3, split_k=split, multitap=mt)
            assert _rel_err(dw, want) < 2e-3, (split, mt)


def test_toy_net_tail_batch_sizes(ext, gemm_flavour):
    """drop_last=False tails of the reference: 304 / 332 / 608 / 664 / 784 rows."""
    from ddpt_b200.ops import gemm
    wk = _bf16(128, 9 * 64, scale=0.05, seed=15)
    for b in (304, 332):
        x = _bf16(b, 12, 12, 64, seed=b)
        y, _ = gemm.conv_fwd(x, wk, 3, 3, out_dtype=torch.float32)
        assert _rel_err(y.permute(0, 3, 1, 2), _conv_ref(x, wk, 128, 64)) < 2e-3


# ---------------------------------------------------------------------------------------------
# amp multi-tensor kernels, fused optimizer
# ---------------------------------------------------------------------------------------------
def test_multi_tensor_scale_and_overflow(ext):
    from ddpt_b200.ops import multi_tensor as mt
    ts = [torch.randn(n, device="cuda") for n in (1, 7, 4096, 5000, 16 * 1024 * 1024 // 4 + 3)]
    outs = [torch.empty_like(t) for t in ts]
    flag = torch.zeros(1, device="cuda")
    scale = torch.tensor([0.125], device="cuda")
    mt.multi_tensor_scale(ts, outs, scale, flag)
    for t, o in zip(ts, outs):
        torch.testing.assert_close(o, t * 0.125, rtol=0, atol=0)
    assert flag.item() == 0
    ts[3][1234] = float("inf")
    mt.multi_tensor_scale(ts, outs, scale, flag)
    assert flag.item() == 1
    many = [torch.randn(33, device="cuda") for _ in range(130)]     # > one launch worth of tensors
    mo = [torch.empty_like(t) for t in many]
    mt.multi_tensor_scale(many, mo, scale, None)
    torch.testing.assert_close(torch.cat(mo), torch.cat(many) * 0.125, rtol=0, atol=0)


def test_multi_tensor_axpby(ext):
    from ddpt_b200.ops import multi_tensor as mt
    xs = [torch.randn(n, device="cuda") for n in (5, 4100)]
    ys = [torch.randn(n, device="cuda") for n in (5, 4100)]
    outs = [torch.empty_like(t) for t in xs]
    flag = torch.zeros(1, device="cuda")
    mt.multi_tensor_axpby(xs, ys, outs, 2.0, -0.5, -1, flag)
    for x, y, o in zip(xs, ys, outs):
        torch.testing.assert_close(o, 2.0 * x - 0.5 * y)
    assert flag.item() == 0
    ys[1][7] = float("nan")
    mt.multi_tensor_axpby(xs, ys, outs, 2.0, -0.5, 0, flag)
    assert flag.item() == 0                                          # only x checked
    mt.multi_tensor_axpby(xs, ys, outs, 2.0, -0.5, 1, flag)
    assert flag.item() == 1


def test_device_loss_scaler_protocol(ext):
    from ddpt_b200.amp import LossScaler
    s = LossScaler(torch.device("cuda"), dynamic=True)
    g = torch.full((1000,), 65536.0, device="cuda")
    s.begin_step(); s.unscale_([g]); s.update()
    assert torch.all(g == 1.0) and s.loss_scale() == 65536.0 and not s.found_inf_host()
    g[5] = float("inf")
    s.begin_step(); s.unscale_([g]); s.update()
    assert s.found_inf_host() and s.loss_scale() == 32768.0 and s.skipped_steps() == 1
    s.load_state_dict({"scale": 4.0, "clean_steps": 1999})
    s.begin_step(); s.unscale_([torch.ones(3, device="cuda")]); s.update()
    assert s.loss_scale() == 8.0


def test_fused_lookahead_sgd_matches_eager(ext):
    """One kernel == torch-style SGD(nesterov) + reference Lookahead, incl. sync steps 1, 11, 21."""
    import torch.nn as nn
    from ddpt_b200 import amp
    from ddpt_b200.ops import fused_optim
    from ddpt_b200.optim import SGD, Lookahead
    torch.manual_seed(0)
    def make():
        return nn.Sequential(nn.Conv2d(3, 8, 3), nn.Flatten(), nn.Linear(8 * 4 * 4, 5)).cuda()
    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    oa = Lookahead(SGD(a.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    ob = Lookahead(SGD(b.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    scaler = amp.LossScaler(torch.device("cuda"))
    assert fused_optim.attach(oa, [a], scaler)
    fused = oa.optimizer._fused
    for step in range(23):
        x = torch.randn(4, 3, 6, 6, device="cuda")
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).square().mean().backward()
        if step == 5:
            oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = 0.05     # scheduler change, no re-capture
        if step == 7:                                                       # overflow: skip incl. counter
            scaler.found_inf_tensor.fill_(1.0)
            before = fused.arena.param_arena.clone()
            oa.step()
            assert torch.equal(before, fused.arena.param_arena)
            scaler.found_inf_tensor.zero_()
            continue
        oa.step(); ob.step()
        for pa, pb in zip(a.parameters(), b.parameters()):
            torch.testing.assert_close(pa, pb, rtol=1e-4, atol=5e-6)
    st = fused.state.tolist()
    assert st[0] == 22 and st[1] == 22 % 10 and st[5] == 1
    # bf16 shadow tracks the master weights
    torch.testing.assert_close(fused.shadow.float(), fused.arena.param_arena, rtol=1e-2, atol=1e-3)
    sd = oa.state_dict()
    assert len(sd["slow_state"]) == 4 and len(sd["fast_state"]) == 4


# ---------------------------------------------------------------------------------------------
# block-scaled fp8 (MXFP8)
# ---------------------------------------------------------------------------------------------
def test_quantize_mxfp8_roundtrip(ext):
    from ddpt_b200.ops import gemm
    x = (torch.randn(64, 256, device="cuda") * torch.logspace(-3, 2, 64, device="cuda").unsqueeze(1)).bfloat16()
    x[3, 32:64] = 0
    q, s = gemm.quantize_mxfp8(x)
    assert q.dtype == torch.uint8 and q.shape == x.shape and s.shape == (64, 8)
    back = gemm.dequantize_mxfp8(q, s)
    blocks = x.float().view(64, 8, 32)
    amax = blocks.abs().amax(-1)
    # scale is the smallest power of two that maps the block into e4m3 range; error <= half an e4m3 ulp of amax
    pow2 = torch.pow(2.0, s.float() - 127.0)
    assert torch.all(amax <= 448.0 * pow2 * 1.0001) and torch.all((amax > 224.0 * pow2) | (amax == 0))
    err = (back.view(64, 8, 32) - blocks).abs().amax(-1)
    assert torch.all(err <= amax * 2.0 ** -4 + 1e-30)
    assert torch.all(back[3, 32:64] == 0)


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (1024, 512, 8192), (304, 256, 512)])
def test_linear_fwd_mxfp8_matches_dequantized_reference(ext, m, n, k):
    from ddpt_b200.ops import gemm
    x = _bf16(m, k, seed=21)
    w = _bf16(n, k, scale=k ** -0.5, seed=22)
    xq, xs = gemm.quantize_mxfp8(x)
    wq, ws = gemm.quantize_mxfp8(w)
    bias = torch.randn(n, device="cuda")
    y, _ = gemm.linear_fwd_mxfp8(xq, xs, wq, ws, bias)
    want = gemm.dequantize_mxfp8(xq, xs) @ gemm.dequantize_mxfp8(wq, ws).t() + bias
    assert _rel_err(y, want) < 6e-3                      # bf16 output rounding only
    assert _rel_err(y, x.float() @ w.float().t() + bias) < 6e-2   # and close to the unquantised product


@pytest.mark.parametrize("n,h,w,ci,co", [(6, 12, 12, 128, 256), (4, 10, 10, 256, 512)])
def test_conv3x3_fwd_mxfp8(ext, n, h, w, ci, co):
    from ddpt_b200.ops import gemm
    x = _bf16(n, h, w, ci, seed=23)
    wk = _bf16(co, 9 * ci, scale=(9 * ci) ** -0.5, seed=24)
    xq, xs = gemm.quantize_mxfp8(x)
    wq, ws = gemm.quantize_mxfp8(wk)
    groups = gemm.stat_row_groups(n, h - 2, w - 2, co, 128)
    stats = torch.zeros(groups * 8 * co, device="cuda")
    y, _ = gemm.conv_fwd_mxfp8(xq, xs, wq, ws, 3, 3, None, False, stats)
    want = _conv_ref(gemm.dequantize_mxfp8(xq, xs), gemm.dequantize_mxfp8(wq, ws), co, ci)
    assert _rel_err(y.permute(0, 3, 1, 2), want) < 6e-3
    part = stats.view(groups * 4, 2, co).sum(0)
    flat = want.permute(0, 2, 3, 1).reshape(-1, co)
    torch.testing.assert_close(part[0], flat.sum(0), rtol=2e-3, atol=5e-2 * math.sqrt(flat.shape[0]))


# ---------------------------------------------------------------------------------------------------------------------
# fp16 operands (kind::f16 with a/b format F16): same kernels, format taken from the tensors' dtype
# ---------------------------------------------------------------------------------------------------------------------
def _f16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.float16)


@pytest.mark.parametrize("n,h,w,ci,co", [(8, 26, 26, 64, 128), (6, 12, 12, 128, 256), (130, 10, 10, 256, 512)])
def test_fp16_conv3x3_fwd_dgrad_wgrad(ext, gemm_flavour, n, h, w, ci, co):
    from ddpt_b200.ops import gemm
    x, wk = _f16(n, h, w, ci, seed=61), _f16(co, 9 * ci, scale=(9 * ci) ** -0.5, seed=62)
    dy = _f16(n, h - 2, w - 2, co, seed=63)
    wt = wk.float().view(co, 3, 3, ci).permute(0, 3, 1, 2)
    y, _ = gemm.conv_fwd(x, wk, 3, 3)
    assert y.dtype == torch.float16
    assert _rel_err(y.permute(0, 3, 1, 2), F.conv2d(x.float().permute(0, 3, 1, 2), wt)) < 2e-3       # fp16: ~1e-3
    dx = gemm.conv_dgrad(dy, wk, 3, 3)
    assert dx.dtype == torch.float16
    assert _rel_err(dx.permute(0, 3, 1, 2), F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt)) < 2e-3
    xg = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wref = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(xg, wref).backward(dy.float().permute(0, 3, 1, 2))
    dw = gemm.conv_wgrad(dy, x, 3, 3)
    assert _rel_err(dw, wref.grad.permute(0, 2, 3, 1)) < 2e-3


def test_fp16_linear_and_splitk(ext):
    from ddpt_b200.ops import gemm
    x, w = _f16(1024, 8192, seed=71), _f16(512, 8192, scale=8192 ** -0.5, seed=72)
    want = x.float() @ w.float().t()
    y, _ = gemm.linear_fwd(x, w, None, block_n=64)
    assert y.dtype == torch.float16 and _rel_err(y, want) < 2e-3
    y2, rows = gemm.linear_fwd_splitk(x, w, None, None, split_k=4, block_n=128)
    assert y2.dtype == torch.float16 and _rel_err(y2, want) < 2e-3


# ---------------------------------------------------------------------------------------------------------------------
# 3x3 data gradient with 64 input channels: filter columns merged into the MMA's N dimension (conv_dgrad_kwmerge.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,h,w,co,pad", [(5, 26, 26, 128, 0),      # Toy_Net conv.3 (one 30-column tile, 7 row bricks)
                                          (3, 9, 37, 64, 0),        # two column tiles, a ragged last one
                                          (2, 4, 30, 128, 0),       # exactly one tile per image
                                          (2, 61, 64, 64, 0),       # three column tiles
                                          (3, 14, 14, 64, 1),       # 'same' convolution (ResNet-style): padding = origin shift
                                          (2, 33, 31, 128, 1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_dgrad_kwmerge_matches_conv_transpose(ext, n, h, w, co, pad, dtype):
    """dx [n, h, w, 64] from dy [n, h - 2 + 2 pad, w - 2 + 2 pad, co]; reference: fp32 conv_transpose2d."""
    from ddpt_b200.ops import gemm
    ci = 64
    ho, wo = h - 2 + 2 * pad, w - 2 + 2 * pad
    g = torch.Generator(device="cuda").manual_seed(81)
    dy = torch.randn(n, ho, wo, co, device="cuda", generator=g).to(dtype)
    wk = (torch.randn(co, 9 * ci, device="cuda", generator=g) * (9 * co) ** -0.5).to(dtype)
    wt = wk.float().view(co, 3, 3, ci).permute(0, 3, 1, 2)
    want = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt, padding=pad)
    out = torch.full((n, h, w, ci), float("nan"), device="cuda", dtype=dtype)
    got = gemm.conv_dgrad(dy, wk, 3, 3, out=out, kwmerge=True, pad=pad)
    assert got.dtype == dtype and torch.isfinite(got.float()).all()       # every output pixel written, none twice-shifted
    assert _rel_err(got.permute(0, 3, 1, 2), want) < (2e-3 if dtype == torch.float16 else 6e-3)
    if pad == 0:        # same answer as the generic kernels
        other = gemm.conv_dgrad(dy, wk, 3, 3, kwmerge=False)
        assert _rel_err(got, other) < 6e-3
    again = gemm.conv_dgrad(dy, wk, 3, 3, kwmerge=True, pad=pad)
    assert torch.equal(got, again)                                        # no atomics: bit-reproducible


def test_conv_dgrad_kwmerge_is_the_default_for_conv3(ext, monkeypatch):
    from ddpt_b200.ops import gemm
    calls = []
    real = ext.conv_gemm
    class Spy:
        def __getattr__(self, name):
            if name == "conv_gemm":
                return lambda *a, **k: (calls.append(a[-1]), real(*a, **k))[1]
            return getattr(ext, name)
    monkeypatch.setattr(gemm, "require_native", lambda: Spy())
    monkeypatch.delenv("DDPT_DGRAD_KWMERGE", raising=False)
    dy, wk = _bf16(4, 24, 24, 128, seed=5), _bf16(128, 576, scale=0.03, seed=6)
    gemm.conv_dgrad(dy, wk, 3, 3)
    assert calls == [4]


# ---------------------------------------------------------------------------------------------------------------------
# conv.0 weight gradient: the CUDA-core kernel (default) and the tensor-core variant (conv0_wgrad_tc.cu, opt-in)
# ---------------------------------------------------------------------------------------------------------------------
def _conv0_wgrad_reference(x_f, dy_f):
    """dW [64, 9], db [64] of Conv2d(1, 64, 3) from x [B, 28, 28] fp32 and dy [B, 26, 26, 64] fp32."""
    xg = x_f.view(-1, 1, 28, 28)
    w = torch.zeros(64, 1, 3, 3, device="cuda", requires_grad=True)
    b = torch.zeros(64, device="cuda", requires_grad=True)
    F.conv2d(xg, w, b).backward(dy_f.permute(0, 3, 1, 2))
    return w.grad.view(64, 9), b.grad


@pytest.mark.parametrize("batch", [3, 333])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("u8", [True, False])
@pytest.mark.parametrize("det", [False, True])
@pytest.mark.parametrize("tc", [True, False, 4])        # tensor cores / CUDA cores 8 channels per thread / 4 per thread
def test_conv0_wgrad_kernels(ext, batch, dtype, u8, det, tc):
    g = torch.Generator(device="cuda").manual_seed(91)
    if u8:
        x = torch.randint(0, 256, (batch, 28, 28), device="cuda", generator=g, dtype=torch.uint8)
        x_f = x.float() / 255.0
    else:
        x = torch.randn(batch, 28, 28, device="cuda", generator=g)
        x_f = x
    dy = (torch.randn(batch, 26, 26, 64, device="cuda", generator=g) * 0.1).to(dtype)
    dw, db = torch.zeros(576, device="cuda"), torch.zeros(64, device="cuda")
    part = torch.zeros(ext.det_workspace_floats(), device="cuda") if det else None
    kw = dict(tensor_core=False, channels_per_thread=4) if tc == 4 else dict(tensor_core=tc, channels_per_thread=8)
    ext.conv0_wgrad(x, dy, dw, db, part, **kw)
    want_w, want_b = _conv0_wgrad_reference(x_f, dy.float())
    # dy is exact (already 16-bit); on the tensor cores x goes in as a 16-bit value + its rounding remainder:
    # fp32-accumulation accuracy either way
    assert _rel_err(dw.view(64, 9), want_w) < 2e-5, _rel_err(dw.view(64, 9), want_w)
    assert _rel_err(db, want_b) < 2e-5
    if det:
        dw2, db2 = torch.zeros(576, device="cuda"), torch.zeros(64, device="cuda")
        ext.conv0_wgrad(x, dy, dw2, db2, part, **kw)
        assert torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("tc", [True, False, 4])
def test_conv0_wgrad_fused_bn_apply(ext, dtype, tc):
    """With the BatchNorm-backward apply folded in: dy = c1 * dz + k0 - k1 * y0, dz = g where relu(bn(y0)) > 0."""
    batch = 200
    g = torch.Generator(device="cuda").manual_seed(92)
    x = torch.randint(0, 256, (batch, 28, 28), device="cuda", generator=g, dtype=torch.uint8)
    y0 = torch.randn(batch, 26, 26, 64, device="cuda", generator=g).to(dtype)
    gin = (torch.randn(batch, 26, 26, 64, device="cuda", generator=g) * 0.05).to(dtype)
    coeffs = torch.empty(4, 64, device="cuda")
    coeffs[0].uniform_(0.5, 1.5, generator=g); coeffs[1].uniform_(-0.5, 0.5, generator=g)          # scale, shift
    coeffs[2].uniform_(-0.2, 0.2, generator=g); coeffs[3].uniform_(0.8, 1.2, generator=g)          # mean, invstd
    bcoef = torch.empty(3, 64, device="cuda")
    bcoef[0].uniform_(0.5, 1.5, generator=g); bcoef[1].uniform_(-0.01, 0.01, generator=g); bcoef[2].uniform_(-0.01, 0.01, generator=g)
    sc, sh, mu, istd = coeffs
    c1, c2, c3 = bcoef
    yf, gf = y0.float(), gin.float()
    dz = torch.where(yf * sc + sh > 0, gf, torch.zeros_like(gf))
    k1 = c1 * istd * c3
    k0 = k1 * mu - c1 * c2
    dy = (c1 * dz + k0 - k1 * yf).to(dtype).float()                 # the kernel rounds dy to the activation type
    want_w, want_b = _conv0_wgrad_reference(x.float() / 255.0, dy)
    dw, db = torch.zeros(576, device="cuda"), torch.zeros(64, device="cuda")
    kw = dict(tensor_core=False, channels_per_thread=4) if tc == 4 else dict(tensor_core=tc, channels_per_thread=8)
    ext.conv0_wgrad(x, torch.empty_like(y0), dw, db, None, y0, gin, coeffs.view(-1), bcoef.view(-1), **kw)
    # fmaf association differs from the torch expression by an ulp of the 16-bit rounding on a few elements
    assert _rel_err(dw.view(64, 9), want_w) < 2e-3, _rel_err(dw.view(64, 9), want_w)
    assert _rel_err(db, want_b) < 2e-3
