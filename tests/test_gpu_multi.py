"""Real multi-GPU tests (NVLink): symmetric memory, fused all-reduce vs NCCL, DDP paths, step engine.

Skipped with fewer than 2 visible devices; run with `gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _need_gpus(n=2):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, fn, args, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    try:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method="env://", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        out = fn(rank, world, *args)
        torch.cuda.synchronize()
        queue.put((rank, "ok", out))
    except BaseException:
        queue.put((rank, "err", traceback.format_exc()))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def run_gpu_ranks(fn, world=2, args=(), timeout=240):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, payload = q.get(timeout=timeout)
            if status != "ok":
                raise AssertionError(f"rank {rank} failed:\n{payload}")
            results[rank] = payload
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    return [results[r] for r in range(world)]


# ------------------------------------------------------------------------------------------------
def _allreduce_vs_nccl(rank, world, provider):
    os.environ["DDPT_SYMM_PROVIDER"] = provider
    from ddpt_b200.ops import require_native
    from ddpt_b200.parallel.symm import SymmetricRegion
    ext = require_native()
    dev = torch.device("cuda", rank)
    n = 64 * 1024 * 1024 // 4
    region = SymmetricRegion(n + 1024, dev)
    info = {"provider": region.provider, "multicast": region.mc_ptr != 0}
    state = torch.tensor([8.0, 0.125, 0.0, 0.0, 0.0], device=dev)
    algos = [1, 2] + ([3] if region.mc_ptr else [])
    for algo in algos:
        for numel in (16 * 1024, 64 * 1024, 1024 * 1024, 16 * 1024 * 1024):
            if algo == 1 and numel > 64 * 1024:
                continue
            g = torch.Generator(device=dev).manual_seed(1000 * rank + numel % 977 + algo)
            src = torch.randn(numel, device=dev, generator=g)
            region.tensor[256:256 + numel].copy_(src)
            want = src.clone()
            dist.all_reduce(want)
            want.mul_(0.125 / world)
            torch.cuda.synchronize(); dist.barrier()
            for it in range(3):                                        # back-to-back: flag reuse
                if it:
                    region.tensor[256:256 + numel].copy_(src)
                    torch.cuda.synchronize(); dist.barrier()
                ext.fused_allreduce(region.peer_ptrs, region.signal_ptrs, region.mc_ptr, rank, 256, numel, state,
                                    state[2:3], 2, algo, 0, 20000)
                torch.cuda.synchronize()
                torch.testing.assert_close(region.tensor[256:256 + numel], want, rtol=2e-5, atol=2e-5)
            assert float(state[2]) == 0.0
    # overflow on one rank reaches everyone
    region.tensor[:4096].fill_(1.0)
    if rank == world - 1:
        region.tensor[77] = float("inf")
    torch.cuda.synchronize(); dist.barrier()
    ext.fused_allreduce(region.peer_ptrs, region.signal_ptrs, region.mc_ptr, rank, 0, 4 * 1024 * 1024, state,
                        state[2:3], 1, 2, 0, 20000)
    torch.cuda.synchronize()
    assert float(state[2]) == 1.0
    return info


@pytest.mark.parametrize("provider", ["torch", "ipc"])
def test_fused_allreduce_matches_nccl(provider):
    _need_gpus(2)
    world = min(torch.cuda.device_count(), 8)
    try:
        infos = run_gpu_ranks(_allreduce_vs_nccl, world, (provider,))
    except AssertionError as e:
        if provider == "torch" and "symmetric" in str(e).lower():
            pytest.skip(f"torch symmetric memory unavailable on this box: {str(e)[-300:]}")
        raise
    assert all(i["provider"] == provider for i in infos)
    print("symmetric memory:", infos[0])


def _ddp_autograd_matches_nccl(rank, world):
    import copy
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as TorchDDP
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    torch.manual_seed(0)
    dev = torch.device("cuda", rank)
    net = nn.Sequential(nn.Conv2d(3, 16, 3), nn.BatchNorm2d(16), nn.ReLU(), nn.Flatten(),
                        nn.Linear(16 * 6 * 6, 64), nn.ReLU(), nn.Linear(64, 10)).to(dev)
    ours_m, theirs_m = copy.deepcopy(net), copy.deepcopy(net)
    ours = DDP(ours_m, device_ids=[rank], comm="fused", bucket_cap_mb=0.01)
    theirs = TorchDDP(theirs_m, device_ids=[rank])
    assert ours.comm == "fused" and len(ours.buckets) >= 2 and getattr(ours.reducer, "native", False)
    g = torch.Generator(device=dev).manual_seed(10 + rank)
    for step in range(4):
        x = torch.randn(16, 3, 8, 8, device=dev, generator=g)
        y = torch.randint(0, 10, (16,), device=dev, generator=g)
        ours.zero_grad(); theirs.zero_grad()
        nn.functional.cross_entropy(ours(x), y).backward()
        nn.functional.cross_entropy(theirs(x), y).backward()
        for a, b in zip(ours_m.parameters(), theirs_m.parameters()):
            torch.testing.assert_close(a.grad, b.grad, rtol=1e-4, atol=1e-6)
    return ours.reducer.backend.provider


def test_ddp_autograd_path_fused_comm_matches_torch_ddp():
    _need_gpus(2)
    run_gpu_ranks(_ddp_autograd_matches_nccl, 2)


def _engine_ranks_identical(rank, world):
    import torch.nn as nn
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    from ddpt_b200.utils import same_seeds
    dev = torch.device("cuda", rank)
    amp.reset()
    same_seeds(7 + rank)                                   # different init per rank: DDP must broadcast rank 0's
    model = Toy_Net().to(dev)
    opt = Lookahead(SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True), k=3, alpha=0.5)
    model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
    ddp = DDP(model, device_ids=[rank], comm="fused")
    loss_fn = nn.CrossEntropyLoss()
    assert ops.attach_step_engine(ddp, loss_fn, opt)
    eng = ddp._ddpt_step_engine
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    ddp.train(); eng.begin_pass()
    losses = []
    for step in range(8):
        x = torch.randint(0, 256, (256, 1, 28, 28), device=dev, generator=g, dtype=torch.uint8)
        y = (x.float().mean((1, 2, 3)) > 127.5).long() + 2 * (x[:, 0, 0, 0] > 127).long()
        eng.run(x, y, True)
    loss_sum, correct, seen = eng.end_pass()
    arena = opt.optimizer._fused.arena.param_arena
    gathered = [torch.empty_like(arena) for _ in range(world)]
    dist.all_gather(gathered, arena)
    assert all(torch.equal(gathered[0], t) for t in gathered), "ranks diverged"
    assert eng.graph_replays >= 6 and loss_sum == loss_sum
    return {"loss": loss_sum / seen, "replays": eng.graph_replays, "provider": ddp.reducer.backend.provider,
            "multicast": ddp.reducer.backend.multicast}


def test_engine_world2_ranks_stay_bit_identical():
    _need_gpus(2)
    out = run_gpu_ranks(_engine_ranks_identical, min(torch.cuda.device_count(), 8))
    print(out[0])
