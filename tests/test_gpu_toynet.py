"""Fused Toy_Net (sm_100a) vs the plain PyTorch fp32 graph: forward, backward, training, engine."""
import copy
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


@pytest.fixture()
def nets():
    from ddpt_b200 import amp
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.ops import toynet_fused
    from ddpt_b200.utils import same_seeds
    amp.reset()
    same_seeds(3)
    ref = Toy_Net().cuda()
    # non-trivial BN affine parameters so their gradients are exercised
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
    fused = copy.deepcopy(ref)
    assert toynet_fused.attach(fused, "bf16")
    return ref, fused


def _batch(b, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(0, 256, (b, 1, 28, 28), device="cuda", generator=g, dtype=torch.uint8)
    y = torch.randint(0, 10, (b,), device="cuda", generator=g)
    return x, y


@pytest.mark.parametrize("batch", [64, 304])
def test_forward_train_and_eval_match_reference(nets, batch):
    ref, fused = nets
    x, _ = _batch(batch)
    ref.train(); fused.train()
    with torch.no_grad():
        want = ref.forward_reference(x)
        got = fused(x)
    assert got.shape == (batch, 10) and _rel(got, want) < 3e-2
    for (k, a), (_, b) in zip(fused.state_dict().items(), ref.state_dict().items()):
        if "running" in k:
            torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-3, msg=k)
        if "num_batches" in k:
            assert int(a) == int(b) == 1
    ref.eval(); fused.eval()
    with torch.no_grad():
        assert _rel(fused(x), ref.forward_reference(x)) < 3e-2


def test_backward_matches_reference(nets):
    """Gradients vs the fp32 graph, judged against what bf16 itself costs: a torch bf16-autocast run of the
    same network is the yardstick (bf16 activations + BatchNorm backward cancellations put both ~10-20 %
    from fp32 at this batch size; see profiles/toynet_grad_error_vs_fp32.txt)."""
    ref, fused = nets
    auto = copy.deepcopy(ref)
    x, y = _batch(128, seed=1)
    ref.train(); fused.train(); auto.train()
    F.cross_entropy(ref.forward_reference(x), y).backward()
    F.cross_entropy(fused(x), y).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = auto.forward_reference(x)
    F.cross_entropy(out.float(), y).backward()
    names = [n for n, _ in ref.named_parameters()]
    bad = {}
    for n, pr, pf, pa in zip(names, ref.parameters(), fused.parameters(), auto.parameters()):
        assert pf.grad is not None and pf.grad.shape == pr.grad.shape, n
        if n.endswith("bias") and n.split(".")[0] + "." + n.split(".")[1] in ("conv.0", "conv.3", "conv.7", "conv.10", "dense.1"):
            # bias followed by BatchNorm: the true gradient is exactly 0, every implementation holds rounding noise
            assert float(pf.grad.abs().max()) < 5e-2, n
            continue
        e_fused, e_auto = _rel(pf.grad, pr.grad), _rel(pa.grad, pr.grad)
        if e_fused > 1.35 * e_auto + 0.01:
            bad[n] = (e_fused, e_auto)
    assert not bad, bad


def test_autograd_training_tracks_fp32_reference(nets):
    """A few SGD steps through model(x)/loss.backward(): losses stay close to the fp32 run."""
    ref, fused = nets
    o_ref = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, nesterov=True)
    o_fus = torch.optim.SGD(fused.parameters(), lr=0.05, momentum=0.9, nesterov=True)
    ref.train(); fused.train()
    lr_, lf_ = [], []
    for step in range(8):
        x, y = _batch(256, seed=100 + step % 2)
        for net, opt, log in ((ref, o_ref, lr_), (fused, o_fus, lf_)):
            opt.zero_grad()
            out = net(x) if net is fused else net.forward_reference(x)
            loss = F.cross_entropy(out, y)
            loss.backward(); opt.step(); log.append(float(loss))
    assert lf_[-1] < lf_[0] * 0.8
    for a, b in zip(lr_, lf_):
        assert abs(a - b) < 0.08 * max(abs(a), 0.1)


def test_step_engine_graph_matches_eager_engine_and_trains():
    """iterate_loader through the CUDA-graph engine: loss falls, Q1 scaling holds, graph == eager."""
    from ddpt_b200 import amp, ops
    from ddpt_b200.data import BatchLoader, synthetic_mnist
    from ddpt_b200.engine import iterate_loader
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.utils import same_seeds
    results = {}
    for use_graph in (False, True):
        amp.reset()
        same_seeds(5)
        model = Toy_Net().cuda()
        opt = Lookahead(SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        loss_fn = nn.CrossEntropyLoss()
        assert ops.attach_step_engine(model, loss_fn, opt, use_graph=use_graph)
        data = synthetic_mnist(2048 + 304, seed=1)
        loader = BatchLoader(data, 256, None, torch.device("cuda"), mode="device")
        loader.quiet = True
        hist = []
        for epoch in range(3):
            model.train()
            hist.append(iterate_loader(loader, model, loss_fn, 0, opt, training=True))
        model.eval()
        with torch.no_grad():
            v = iterate_loader(loader, model, loss_fn, 0, None, training=False)
        eng = model._ddpt_step_engine
        results[use_graph] = (hist, v, eng.graph_replays)
        assert hist[-1][0] < hist[0][0] * 0.7, hist                  # learns
        assert 0.5 < hist[0][0] * 256 < 3.0                          # Q1: logged loss is CE / B
        assert hist[0][2] == 0.1 and 0 <= v[1] <= 1
        st = amp.current_state().scaler.state_dict()
        assert st["scale"] == 65536.0 and st["skipped"] == 0
    assert results[True][2] > 0 and results[False][2] == 0
    for (a, b) in zip(results[True][0], results[False][0]):
        assert abs(a[0] - b[0]) < 0.05 * abs(b[0]) and abs(a[1] - b[1]) < 0.05
    amp.reset()


def test_engine_overflow_skips_step_and_halves_scale():
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    amp.reset()
    model = Toy_Net().cuda()
    opt = Lookahead(SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
    loss_fn = nn.CrossEntropyLoss()
    assert ops.attach_step_engine(model, loss_fn, opt)
    eng = model._ddpt_step_engine
    x, y = _batch(64)
    model.train()
    eng.begin_pass()
    for _ in range(3):
        eng.run(x, y, True)                                          # eager, capture, replay
    fo = opt.optimizer._fused
    before = fo.arena.param_arena.clone()
    with torch.no_grad():
        model.dense[4].bias.fill_(float("nan"))                      # poisons logits -> inf/nan gradients
    poisoned = fo.arena.param_arena.clone()
    eng.run(x, y, True)
    assert torch.equal(torch.nan_to_num(fo.arena.param_arena), torch.nan_to_num(poisoned))   # step skipped
    sd = amp.current_state().scaler.state_dict()
    assert sd["scale"] == 32768.0 and sd["skipped"] == 1
    assert fo.state.tolist()[5] == 1 and fo.state.tolist()[0] == 3
    amp.reset()


def test_fp8_block_scaled_forward_path_trains():
    """opt_level O1-fp8: conv.7 / conv.10 / dense.1 forward on MXFP8 tensor cores; loss still falls."""
    from ddpt_b200 import amp, ops
    from ddpt_b200.data import BatchLoader, synthetic_mnist
    from ddpt_b200.engine import iterate_loader
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.utils import same_seeds
    amp.reset()
    same_seeds(5)
    model = Toy_Net().cuda()
    ref = copy.deepcopy(model)
    opt = Lookahead(SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    model, opt = amp.initialize(model, optimizers=opt, opt_level="O1-fp8")
    assert model._fused.compute == "fp8"
    x, _ = _batch(128, seed=9)
    model.train(); ref.train()
    with torch.no_grad():
        assert _rel(model(x), ref.forward_reference(x)) < 0.2       # e4m3 (3 mantissa bits) operands on three layers
    loss_fn = nn.CrossEntropyLoss()
    assert ops.attach_step_engine(model, loss_fn, opt)
    data = synthetic_mnist(2048, seed=1)
    loader = BatchLoader(data, 256, None, torch.device("cuda"), mode="device")
    loader.quiet = True
    hist = [iterate_loader(loader, model, loss_fn, 0, opt, training=True)[0] for _ in range(3)]
    assert hist[-1] < hist[0] * 0.75, hist
    amp.reset()


def test_chw_to_hwc_and_gemm_dispatch():
    """The flatten-order -> channels-last transpose ahead of conv.10's BatchNorm backward, and which GEMM kernel each
    Toy_Net layer is routed to by default (persistent / halo reuse / CTA pair)."""
    from ddpt_b200 import ops
    from ddpt_b200.ops import gemm
    ext = ops.require_native()
    for n, c, p in ((3, 512, 16), (1024, 512, 16), (5, 64, 8)):
        x = torch.randn(n, c * p, device="cuda").bfloat16()
        out = torch.empty(n, p, c, device="cuda", dtype=torch.bfloat16)
        ext.chw_to_hwc(x, out, c, p)
        assert torch.equal(out, x.view(n, c, p).transpose(1, 2).contiguous())
    env = {k: os.environ.pop(k, None) for k in ("DDPT_GEMM_PAIR", "DDPT_GEMM_HALO", "DDPT_GEMM_PERSISTENT")}
    try:
        B = 1024
        mode = lambda nhw, ck, n, bn: gemm._gemm_mode(None, None, nhw, 3, 3, ck, n, bn)[0]
        assert mode((B, 24, 24), 64, 128, 128) == 2       # conv.3 forward: resident weights, halo reuse
        assert mode((B, 26, 26), 128, 64, 64) == 2        # conv.3 data gradient: halo reuse
        assert mode((B, 10, 10), 128, 256, 256) == 3      # conv.7 forward: CTA pair
        assert mode((B, 8, 8), 256, 512, 256) == 3        # conv.10 forward: CTA pair
        assert mode((B, 12, 12), 256, 128, 128) == 3      # conv.7 data gradient: CTA pair
    finally:
        for k, v in env.items():
            if v is not None:
                os.environ[k] = v


def _train_steps(steps=20, batch=256, seed=3):
    import torch.nn as nn
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.utils import same_seeds
    amp.reset()
    same_seeds(seed)
    model = Toy_Net().cuda()
    opt = Lookahead(SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True), k=5, alpha=0.5)
    model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
    loss_fn = nn.CrossEntropyLoss()
    assert ops.attach_step_engine(model, loss_fn, opt)
    eng = model._ddpt_step_engine
    g = torch.Generator(device="cuda").manual_seed(seed)
    model.train(); eng.begin_pass()
    for _ in range(steps):
        x = torch.randint(0, 256, (batch, 1, 28, 28), device="cuda", generator=g, dtype=torch.uint8)
        y = (x.float().mean((1, 2, 3)) > 127.5).long() + 2 * (x[:, 0, 0, 0] > 127).long()
        eng.run(x, y, True)
    loss_sum, correct, seen = eng.end_pass()
    params = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).clone()
    bufs = torch.cat([b.detach().double().reshape(-1) for b in model.buffers()]).clone()
    amp.reset()
    return params, bufs, loss_sum


def test_deterministic_mode_is_bit_reproducible(monkeypatch):
    """DDPT_DETERMINISTIC=1 (reference: cudnn.deterministic, utils/customized.py:12-13): 20 optimizer steps twice ->
    identical parameters, BatchNorm buffers and accumulated loss, bit for bit; and the mode trains like the fast one."""
    monkeypatch.setenv("DDPT_DETERMINISTIC", "1")
    p1, b1, l1 = _train_steps()
    p2, b2, l2 = _train_steps()
    assert torch.equal(p1.view(torch.int32), p2.view(torch.int32)), float((p1 - p2).abs().max())
    assert torch.equal(b1, b2) and l1 == l2
    monkeypatch.setenv("DDPT_DETERMINISTIC", "0")
    p3, b3, l3 = _train_steps()
    assert float((p1 - p3).abs().max()) < 2e-2 * float(p1.abs().max())      # same trajectory up to rounding
    assert l3 == pytest.approx(l1, rel=2e-2)


# ---------------------------------------------------------------------------------------------------------------------
# native fp16 path ("O1-fp16": Apex O1's precision in the reference, utils/train.py:58)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture()
def nets_fp16():
    from ddpt_b200 import amp
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.ops import toynet_fused
    from ddpt_b200.utils import same_seeds
    amp.reset()
    same_seeds(3)
    ref = Toy_Net().cuda()
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
    fused = copy.deepcopy(ref)
    assert toynet_fused.attach(fused, "fp16")
    return ref, fused


@pytest.mark.parametrize("batch", [64, 304])
def test_fp16_forward_train_and_eval_match_reference(nets_fp16, batch):
    ref, fused = nets_fp16
    x, _ = _batch(batch)
    ref.train(); fused.train()
    with torch.no_grad():
        want = ref.forward_reference(x)
        got = fused(x)
    assert fused._fused.act_dtype == torch.float16 and fused._fused.workspace(batch).y[1].dtype == torch.float16
    assert _rel(got, want) < 1e-2                      # fp16 has 3 more mantissa bits than bf16 (bf16 bound: 3e-2)
    for (k, a), (_, b) in zip(fused.state_dict().items(), ref.state_dict().items()):
        if "running" in k:
            torch.testing.assert_close(a, b, rtol=5e-3, atol=1e-3, msg=k)
    ref.eval(); fused.eval()
    with torch.no_grad():
        assert _rel(fused(x), ref.forward_reference(x)) < 1e-2


def test_fp16_backward_matches_reference(nets_fp16):
    """Same yardstick as the bf16 test: a torch fp16-autocast run of the network (BatchNorm-backward cancellations
    dominate the distance from fp32, not the storage format)."""
    ref, fused = nets_fp16
    auto = copy.deepcopy(ref)
    x, y = _batch(128, seed=1)
    ref.train(); fused.train(); auto.train()
    # fp16 gradients need the loss scale (the reference's Apex O1 does exactly this): scale the loss, unscale the result
    scale = 1024.0
    F.cross_entropy(ref.forward_reference(x), y).backward()
    (F.cross_entropy(fused(x), y) * scale).backward()
    with torch.autocast("cuda", dtype=torch.float16):
        out = auto.forward_reference(x)
    (F.cross_entropy(out.float(), y) * scale).backward()
    names = [n for n, _ in ref.named_parameters()]
    bad = {}
    for n, pr, pf, pa in zip(names, ref.parameters(), fused.parameters(), auto.parameters()):
        g, ga = pf.grad / scale, pa.grad / scale
        assert torch.isfinite(g).all(), n
        if n.endswith("bias") and n.split(".")[0] + "." + n.split(".")[1] in ("conv.0", "conv.3", "conv.7", "conv.10", "dense.1"):
            assert float(g.abs().max()) < 5e-2, n
            continue
        e_fused, e_auto = _rel(g, pr.grad), _rel(ga, pr.grad)
        if e_fused > 1.35 * e_auto + 0.01:
            bad[n] = (e_fused, e_auto)
    assert not bad, bad


def test_fp16_step_engine_trains_with_dynamic_loss_scale():
    """`--opt_level O1-fp16` end to end: fused fp16 kernels + CUDA-graph engine + dynamic loss scaling; trains like bf16."""
    import torch.nn as nn2
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.utils import same_seeds
    losses = {}
    for level in ("O1-fp16", "O1"):
        amp.reset(); same_seeds(5)
        model = Toy_Net().cuda()
        opt = Lookahead(SGD(model.parameters(), lr=0.5, momentum=0.9, nesterov=True), k=5, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level=level)
        assert model._fused.act_dtype == (torch.float16 if level == "O1-fp16" else torch.bfloat16)
        assert opt.optimizer._fused.shadow.dtype == model._fused.act_dtype
        assert ops.attach_step_engine(model, nn2.CrossEntropyLoss(), opt)
        eng = model._ddpt_step_engine
        g = torch.Generator(device="cuda").manual_seed(5)
        model.train()
        out = []
        for it in range(24):
            x = torch.randint(0, 256, (256, 1, 28, 28), device="cuda", generator=g, dtype=torch.uint8)
            y = (x.float().mean((1, 2, 3)) > 127.5).long() + 2 * (x[:, 0, 0, 0] > 127).long()
            eng.begin_pass(); eng.run(x, y, True)
            loss_sum, _, seen = eng.end_pass()
            out.append(loss_sum / seen * 256)
        losses[level] = out
        scale = amp.current_state().scaler.state.tolist()
        assert scale[0] >= 1.0 and eng.graph_replays >= 20
        amp.reset()
    assert losses["O1-fp16"][-1] < 0.75 * losses["O1-fp16"][0], losses["O1-fp16"]
    for a, b in zip(losses["O1-fp16"][:6], losses["O1"][:6]):
        assert a == pytest.approx(b, rel=0.05, abs=0.02)
