"""C4/C9/C10/C11: seeds, schedulers (incl. quirks Q2/Q3), early stopping."""
import math

import pytest
import torch

from ddpt_b200.optim import SGD, LambdaLR, ReduceLROnPlateau, warmup_lambda
from ddpt_b200.utils import EarlyStopping, same_seeds


def test_same_seeds_reproducible():
    same_seeds(42)
    a = (torch.rand(3), __import__("random").random(), __import__("numpy").random.rand())
    same_seeds(42)
    b = (torch.rand(3), __import__("random").random(), __import__("numpy").random.rand())
    assert torch.equal(a[0], b[0]) and a[1:] == b[1:]
    assert torch.backends.cudnn.deterministic and not torch.backends.cudnn.benchmark


def test_early_stopping_truth_table():
    es = EarlyStopping(patience=3, verbose=False)
    for loss in [1.0, 0.9, 0.9]:       # equal score counts as improvement when delta == 0
        es(loss)
    assert es.counter == 0 and not es.early_stop
    es(0.95); es(0.95)
    assert es.counter == 2 and not es.early_stop
    es(0.8)
    assert es.counter == 0
    for _ in range(3):
        es(0.85)
    assert es.early_stop
    acc = EarlyStopping(patience=1, verbose=False)
    acc(0.5, loss=False); acc(0.4, loss=False)
    assert acc.early_stop


def test_early_stopping_prints_and_roundtrips(capsys):
    es = EarlyStopping(patience=1, verbose=True)
    es(1.0); es(2.0)
    assert "EarlyStopping" in capsys.readouterr().out
    other = EarlyStopping(patience=1)
    other.load_state_dict(es.state_dict())
    assert other.early_stop and other.best_score == es.best_score


def _lr_series(kind, epochs=500, warm=10, lr=0.1, always_step=False):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = SGD([p], lr=lr, momentum=0.9, nesterov=True)
    sched = LambdaLR(opt, warmup_lambda(kind, warm, epochs))
    out = []
    for e in range(30):
        out.append(opt.param_groups[0]["lr"])
        if e <= warm or always_step:
            sched.step()
    return out


def test_linear_warmup_series_matches_logged_lr():
    s = _lr_series("linear")
    assert s[0] == 0.0                                   # Q3: epoch 0 at lr 0
    for e in range(11):
        assert s[e] == pytest.approx(0.1 * e / 10)
    assert all(v == pytest.approx(0.1) for v in s[11:])


def test_cosine_never_decays_q2():
    s = _lr_series("cosine")
    frozen = 0.1 * 0.5 * (math.cos((11 - 10) / (500 - 10) * math.pi) + 1)
    assert frozen == pytest.approx(0.09999897, abs=1e-8)   # value in the reference's TB log
    assert all(v == pytest.approx(frozen) for v in s[11:])
    fixed = _lr_series("cosine", always_step=True)
    assert fixed[29] < fixed[12] < 0.1


def test_lambda_lr_matches_torch():
    p = torch.nn.Parameter(torch.zeros(1))
    a, b = SGD([p], lr=0.3, momentum=0.9), torch.optim.SGD([p], lr=0.3, momentum=0.9)
    fn = warmup_lambda("cosine", 4, 20)
    ours, theirs = LambdaLR(a, fn), torch.optim.lr_scheduler.LambdaLR(b, fn)
    for _ in range(20):
        assert a.param_groups[0]["lr"] == pytest.approx(b.param_groups[0]["lr"], rel=1e-12)
        ours.step(); b.step(); theirs.step()


def test_plateau_matches_torch():
    g = torch.Generator().manual_seed(0)
    metrics = (1.0 + 0.01 * torch.randn(80, generator=g)).tolist()
    p = torch.nn.Parameter(torch.zeros(1))
    a, b = SGD([p], lr=0.1), torch.optim.SGD([p], lr=0.1)
    ours = ReduceLROnPlateau(a, mode="min", factor=0.1, patience=6)
    theirs = torch.optim.lr_scheduler.ReduceLROnPlateau(b, mode="min", factor=0.1, patience=6)
    for m in metrics:
        ours.step(m); theirs.step(m)
        assert a.param_groups[0]["lr"] == pytest.approx(b.param_groups[0]["lr"], rel=1e-12)
    assert a.param_groups[0]["lr"] < 0.1


def test_plateau_verbose_message(capsys):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = SGD([p], lr=0.1)
    s = ReduceLROnPlateau(opt, patience=0, verbose=True)
    s.step(1.0); s.step(2.0)
    assert "reducing learning rate of group 0 to 1.0000e-02" in capsys.readouterr().out
