"""L3-L5 end to end on CPU: iterate_loader semantics (Q1), launcher run, resume, epoch-0 oracle."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn as nn

from conftest import ROOT
from ddpt_b200 import amp
from ddpt_b200.config import parse_args
from ddpt_b200.data import BatchLoader, ShardedSampler, load_mnist, synthetic_mnist
from ddpt_b200.engine import iterate_loader, train
from ddpt_b200.models import Toy_Net
from ddpt_b200.optim import SGD, Lookahead
from ddpt_b200.utils import same_seeds


def test_iterate_loader_q1_double_normalisation_and_signature():
    amp.reset()
    same_seeds(0)
    model = Toy_Net()
    data = synthetic_mnist(96)
    loader = BatchLoader(data, 32, None, mode="torch")
    loader.quiet = True
    loss_fn = nn.CrossEntropyLoss()
    model.eval()
    with torch.no_grad():
        loss, acc = iterate_loader(loader, model, loss_fn, 0, None, training=False)
        want = sum(float(loss_fn(model(x), y)) / 32 * 32 for x, y in loader) / 96
    assert loss == pytest.approx(want, rel=1e-5) and 0 <= acc <= 1
    assert 0.05 < loss * 32 < 5                     # logged loss is CE / B
    opt = Lookahead(SGD(model.parameters(), lr=0.0, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    model.train()
    out = iterate_loader(loader, model, loss_fn, 0, opt, training=True)
    assert len(out) == 3 and out[2] == 0.0


def test_train_two_epochs_writes_reference_artifacts(tmp_path):
    args = parse_args(["-n", "unit", "--synthetic", "-e", "2", "-b", "64", "--max_steps", "2", "-w", "1",
                       "-p", str(tmp_path), "--quiet", "--save_resume"])
    out = train(args)
    assert [h["lr"] for h in out["history"]] == [0.0, pytest.approx(0.1)]
    assert (tmp_path / "unit.pt").exists() and (tmp_path / "logs" / "unit").is_dir()
    state = torch.load(tmp_path / "unit.pt", weights_only=True)
    assert len(state) == 37 and not any(k.startswith("module.") for k in state)
    # resume continues at epoch 2 with restored optimizer/scheduler state
    args2 = parse_args(["-n", "unit", "--synthetic", "-e", "3", "-b", "64", "--max_steps", "2", "-w", "1",
                        "-p", str(tmp_path), "--quiet", "--resume", str(tmp_path / "unit.resume.pt")])
    out2 = train(args2)
    assert [h["epoch"] for h in out2["history"]] == [2]
    assert out2["history"][0]["lr"] == pytest.approx(0.1)


def test_launcher_world2_gloo(tmp_path):
    """`python -m torch.distributed.run --nproc-per-node 2 run.py ...` on CPU (config 1 plumbing)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "run.py"),
           "-n=DDP_warmup", "--synthetic", "-e", "2", "-b", "32", "--max_steps", "2", "-p", str(tmp_path),
           "--quiet"]
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "Running DDP on rank: 0" in res.stdout and "Running DDP on rank: 1" in res.stdout
    assert res.stdout.count("epoch: 001/2") >= 1
    assert (tmp_path / "DDP_warmup.pt").exists()


@pytest.mark.slow
def test_epoch0_oracle_from_shipped_logs(reference_root):
    """SURVEY.md section 4 oracle: seed 42, world 2, rank 0, B 1024, lr 0 in epoch 0.

    Logged by the reference: train_loss 0.002334, valid_loss 0.002310, valid_acc 0.1075."""
    amp.reset()
    root = os.path.join(reference_root, "datasets")
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    train_data, valid_data = load_mnist(root, True), load_mnist(root, False)
    same_seeds(42)
    tl = BatchLoader(train_data, 1024, ShardedSampler(60000, num_replicas=2, rank=0), mode="torch")
    vl = BatchLoader(valid_data, 1024, None, mode="torch")
    tl.quiet = vl.quiet = True
    same_seeds(42)
    model = Toy_Net()
    opt = Lookahead(SGD(model.parameters(), lr=0.0, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    loss_fn = nn.CrossEntropyLoss()
    model.train()
    with torch.no_grad():                                        # lr == 0: weights cannot move; skip backward
        tloss, tacc = iterate_loader(tl, model, loss_fn, 0, None, training=False)
    model.eval()
    with torch.no_grad():
        vloss, vacc = iterate_loader(vl, model, loss_fn, 0, None, training=False)
    assert tloss == pytest.approx(0.002334, abs=2e-6)
    assert vloss == pytest.approx(0.002310, abs=2e-6)
    assert tacc == pytest.approx(0.093733, abs=3e-4) and vacc == pytest.approx(0.1075, abs=3e-4)
    assert int(model.conv[1].num_batches_tracked) == 30
