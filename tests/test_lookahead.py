"""C6/C7: Nesterov SGD and Lookahead semantics (reference utils/lookahead.py, quirk Q7)."""
import copy

import numpy as np
import pytest
import torch

from conftest import load_reference_module
from ddpt_b200.optim import SGD, Lookahead


def _numpy_lookahead_sgd(w0, grads, lr, mu, k, alpha):
    """Independent re-derivation: nesterov SGD + lookahead with lazy slow init."""
    w, buf, slow, counter, sync_steps = w0.copy(), None, None, 0, []
    for step, g in enumerate(grads, 1):
        buf = g.copy() if buf is None else mu * buf + g
        w = w - lr * (g + mu * buf)
        if counter == 0:
            if slow is None:
                slow = w.copy()
            slow = slow + alpha * (w - slow)
            w = slow.copy()
            sync_steps.append(step)
        counter = (counter + 1) % k
    return w, sync_steps


def test_lookahead_matches_numpy_derivation():
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(7, dtype=torch.float64))
    w0 = p.detach().numpy().copy()
    grads = [np.random.RandomState(i).randn(7) for i in range(25)]
    opt = Lookahead(SGD([p], lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    snapshots = []
    for g in grads:
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        snapshots.append(p.detach().numpy().copy())
    want, sync = _numpy_lookahead_sgd(w0, grads, 0.1, 0.9, 10, 0.5)
    assert sync == [1, 11, 21]                              # Q7: steps 1, 11, 21
    np.testing.assert_allclose(snapshots[-1], want, rtol=1e-12)
    # step-1 sync is the identity: slow initialised from post-step fast weights
    buf = grads[0]
    np.testing.assert_allclose(snapshots[0], w0 - 0.1 * (buf + 0.9 * buf), rtol=1e-12)


def test_sgd_matches_torch_sgd():
    torch.manual_seed(1)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(4))]
    b = [torch.nn.Parameter(t.detach().clone()) for t in a]
    ours = SGD(a, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
    theirs = torch.optim.SGD(b, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
    for i in range(6):
        for x, y in zip(a, b):
            g = torch.randn_like(x)
            x.grad, y.grad = g.clone(), g.clone()
        ours.step(); theirs.step()
    for x, y in zip(a, b):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-7)


def test_shares_param_groups_and_is_real_optimizer():
    p = torch.nn.Parameter(torch.zeros(3))
    inner = SGD([p], lr=0.1, momentum=0.9, nesterov=True)
    la = Lookahead(inner, k=10, alpha=0.5)
    assert la.param_groups is inner.param_groups and la.param_groups[0]["counter"] == 0
    inner.param_groups[0]["lr"] = 0.5
    assert la.param_groups[0]["lr"] == 0.5
    p.grad = torch.ones(3)
    la.zero_grad()                                           # reference raises AttributeError on torch 2.x
    assert p.grad is None
    assert isinstance(la, torch.optim.Optimizer)


def test_state_dict_roundtrip_continues_identically():
    torch.manual_seed(2)
    def make():
        p = torch.nn.Parameter(torch.arange(6, dtype=torch.float32))
        return p, Lookahead(SGD([p], lr=0.1, momentum=0.9, nesterov=True), k=3, alpha=0.5)
    p1, o1 = make()
    grads = [torch.randn(6) for _ in range(9)]
    for g in grads[:4]:
        p1.grad = g.clone(); o1.step()
    state = copy.deepcopy(o1.state_dict())
    assert set(state) >= {"fast_state", "slow_state", "param_groups"}
    p2, o2 = make()
    with torch.no_grad():
        p2.copy_(p1)
    o2.load_state_dict(state)
    for g in grads[4:]:
        p1.grad = g.clone(); o1.step()
        p2.grad = g.clone(); o2.step()
    torch.testing.assert_close(p1, p2, rtol=0, atol=0)


def test_matches_reference_lookahead(reference_root):
    ref = load_reference_module("lookahead")
    torch.manual_seed(3)
    w = torch.randn(4, 4)
    pa, pb = torch.nn.Parameter(w.clone()), torch.nn.Parameter(w.clone())
    ours = Lookahead(SGD([pa], lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    theirs = ref.Lookahead(torch.optim.SGD([pb], lr=0.1, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    for i in range(23):
        g = torch.randn(4, 4)
        pa.grad, pb.grad = g.clone(), g.clone()
        ours.step(); theirs.step()
        torch.testing.assert_close(pa, pb, rtol=1e-6, atol=1e-7)


def test_bad_hyperparameters():
    p = torch.nn.Parameter(torch.zeros(1))
    with pytest.raises(ValueError):
        Lookahead(SGD([p], lr=0.1), k=0)
    with pytest.raises(ValueError):
        Lookahead(SGD([p], lr=0.1), alpha=1.5)
