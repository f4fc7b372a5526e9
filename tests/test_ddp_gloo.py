"""C13 / N2-N5 on CPU: our DDP wrapper over gloo, world_size 2 (BASELINE.json config 1)."""
import torch
import torch.nn as nn

from mp_utils import run_ranks


def _small_net():
    return nn.Sequential(nn.Conv2d(1, 4, 3), nn.BatchNorm2d(4), nn.ReLU(), nn.Flatten(),
                         nn.Linear(4 * 6 * 6, 8), nn.BatchNorm1d(8), nn.ReLU(), nn.Linear(8, 3))


def _grads_equal_average(rank, world):
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    torch.manual_seed(100 + rank)                   # different init per rank: ctor must broadcast rank 0's
    model = _small_net()
    ddp = DDP(model, bucket_cap_mb=0.0001)           # force several buckets
    assert len(ddp.buckets) > 2
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], g) for g in gathered)

    # reference gradients: same weights, plain module, every rank's batch, averaged
    import copy
    plain = copy.deepcopy(model)
    xs = [torch.randn(6, 1, 8, 8, generator=torch.Generator().manual_seed(r)) for r in range(world)]
    ys = [torch.randint(0, 3, (6,), generator=torch.Generator().manual_seed(50 + r)) for r in range(world)]
    want = [torch.zeros_like(p) for p in plain.parameters()]
    for r in range(world):
        plain.zero_grad()
        nn.functional.cross_entropy(plain(xs[r]), ys[r]).backward()
        for w, p in zip(want, plain.parameters()):
            w += p.grad / world
    ddp.zero_grad()
    nn.functional.cross_entropy(ddp(xs[rank]), ys[rank]).backward()
    for w, p in zip(want, model.parameters()):
        torch.testing.assert_close(p.grad, w, rtol=1e-5, atol=1e-6)
    assert ddp.reducer.stats["launches"] == len(ddp.buckets)
    import os
    if os.environ.get("DDPT_PY_REDUCER") != "1":
        assert getattr(ddp.reducer, "native", False), "the C++ reducer should be active when the extension is built"

    # no_sync: local gradient only, next synced backward folds it in
    ddp.zero_grad()
    with ddp.no_sync():
        nn.functional.cross_entropy(ddp(xs[rank]), ys[rank]).backward()
    local = [p.grad.clone() for p in model.parameters()]
    nn.functional.cross_entropy(ddp(xs[rank]), ys[rank]).backward()
    return True


def test_ddp_gradients_are_rank_average():
    assert run_ranks(_grads_equal_average, 2) == [True, True]


def _train_steps_match_torch_ddp(rank, world):
    """Weights stay bit-identical across ranks and track torch's DDP numerically."""
    import copy
    from torch.nn.parallel import DistributedDataParallel as TorchDDP
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    torch.manual_seed(7)
    base = _small_net()
    ours_m, theirs_m = copy.deepcopy(base), copy.deepcopy(base)
    ours = DDP(ours_m)
    theirs = TorchDDP(theirs_m)
    o1 = Lookahead(SGD(ours_m.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=3, alpha=0.5)
    o2 = torch.optim.SGD(theirs_m.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    from conftest import REFERENCE
    g = torch.Generator().manual_seed(1234 + rank)
    slow = None
    for step in range(7):
        x, y = torch.randn(8, 1, 8, 8, generator=g), torch.randint(0, 3, (8,), generator=g)
        o1.zero_grad(); o2.zero_grad()
        nn.functional.cross_entropy(ours(x), y).backward()
        nn.functional.cross_entropy(theirs(x), y).backward()
        o1.step(); o2.step()
        # emulate lookahead(k=3, alpha=.5) on the torch side
        if step % 3 == 0:
            with torch.no_grad():
                if slow is None:
                    slow = [p.detach().clone() for p in theirs_m.parameters()]
                for s, p in zip(slow, theirs_m.parameters()):
                    s += 0.5 * (p - s); p.copy_(s)
    for a, b in zip(ours_m.parameters(), theirs_m.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    for (ka, a), (kb, b) in zip(ours_m.state_dict().items(), theirs_m.state_dict().items()):
        if "running" in ka:
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    flat = torch.cat([p.detach().reshape(-1) for p in ours_m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], t) for t in gathered), "ranks diverged"
    assert list(ours.state_dict())[0].startswith("module.")
    return True


def test_training_tracks_torch_ddp_and_ranks_stay_identical():
    assert run_ranks(_train_steps_match_torch_ddp, 2) == [True, True]


def _buffers_follow_rank0(rank, world):
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    torch.manual_seed(0)
    model = _small_net()
    ddp = DDP(model)
    ddp.train()
    x = torch.randn(8, 1, 8, 8, generator=torch.Generator().manual_seed(rank)) * (1 + 5 * rank)
    ddp(x)                      # BN stats now differ per rank
    ddp(x)                      # grad-enabled forward re-broadcasts rank 0's buffers first
    ddp.eval()
    with torch.no_grad():
        ddp(x)                  # first no-grad forward after training also syncs
    mean = model[1].running_mean.clone()
    gathered = [torch.empty_like(mean) for _ in range(world)]
    torch.distributed.all_gather(gathered, mean)
    assert torch.equal(gathered[0], gathered[1])
    return True


def test_bn_buffers_broadcast_from_rank0():
    assert run_ranks(_buffers_follow_rank0, 2) == [True, True]


def _shape_mismatch(rank, world):
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    model = nn.Linear(4, 3 + rank)
    try:
        DDP(model)
    except RuntimeError as e:
        return "shapes differ" in str(e)
    return False


def test_shape_verification():
    assert run_ranks(_shape_mismatch, 2) == [True, True]
