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
    info = {"provider": region.provider, "multicast": region.mc_ptr != 0, "checked": []}
    state = torch.tensor([8.0, 0.125, 0.0, 0.0, 0.0], device=dev)

    def call(offset, numel, channel, algo, wire):
        ext.fused_allreduce(region.peer_ptrs, region.signal_ptrs, region.mc_ptr, region.scratch_ptrs,
                            region.mc_scratch_ptr, rank, offset, numel, state, state[2:3], channel, algo, wire, 0, 20000)

    algos = [1, 2] + ([3] if region.mc_ptr else [])
    for wire in (0, 1):
        for algo in algos:
            for numel in (16 * 1024, 64 * 1024 - 64 * 8, 1024 * 1024, 16 * 1024 * 1024):
                if algo == 1 and numel > 64 * 1024:
                    continue
                g = torch.Generator(device=dev).manual_seed(1000 * rank + numel % 977 + algo)
                src = torch.randn(numel, device=dev, generator=g)
                want = src.bfloat16().float() if wire else src.clone()
                dist.all_reduce(want)
                want.mul_(0.125 / world)
                for it in range(3):                                        # back-to-back: epochs / slot parity
                    region.tensor[256:256 + numel].copy_(src)
                    torch.cuda.synchronize(); dist.barrier()
                    call(256, numel, 2, algo, wire)
                    torch.cuda.synchronize()
                    got = region.tensor[256:256 + numel]
                    if wire and algo != 1:
                        torch.testing.assert_close(got, want, rtol=2 ** -7, atol=1e-6)
                    else:
                        torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)
                    # bit-identical across ranks
                    mine = got.view(torch.int32).clone()
                    ref = mine.clone()
                    dist.broadcast(ref, 0)
                    assert torch.equal(mine, ref), f"rank {rank} differs from rank 0 (algo {algo}, wire {wire})"
                assert float(state[2]) == 0.0
                info["checked"].append((wire, algo, numel))
    # overflow on one rank reaches everyone (large bucket, many CTAs), then clears
    for wire in (0, 1):
        region.tensor[:4 * 1024 * 1024].fill_(1.0)
        if rank == world - 1:
            region.tensor[4 * 1024 * 1024 - 3] = float("inf")
        torch.cuda.synchronize(); dist.barrier()
        call(0, 4 * 1024 * 1024, 1, 0, wire)
        torch.cuda.synchronize()
        assert float(state[2]) == 1.0
        state[2] = 0.0
        region.tensor[:4 * 1024 * 1024].fill_(1.0)
        torch.cuda.synchronize(); dist.barrier()
        call(0, 4 * 1024 * 1024, 1, 0, wire)
        torch.cuda.synchronize()
        assert float(state[2]) == 0.0
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
    print("symmetric memory:", {k: v for k, v in infos[0].items() if k != "checked"}, len(infos[0]["checked"]), "cases")


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


def _engine_pg_matches_fused(rank, world):
    """with comm='pg' the step engine must not all-reduce bucket 0 while the side-stream weight
    gradient is still being written.  Parameters after N steps must equal the comm='fused' run bit for bit on
    the fp32 wire (both are exact fp32 sums / world at world 2)."""
    import torch.nn as nn
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    from ddpt_b200.utils import same_seeds
    dev = torch.device("cuda", rank)
    results = {}
    for comm in ("pg", "fused"):
        amp.reset()
        same_seeds(11)
        model = Toy_Net().to(dev)
        opt = Lookahead(SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True), k=3, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        ddp = DDP(model, device_ids=[rank], comm=comm)
        assert ddp.comm == comm
        loss_fn = nn.CrossEntropyLoss()
        assert ops.attach_step_engine(ddp, loss_fn, opt)
        eng = ddp._ddpt_step_engine
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        ddp.train(); eng.begin_pass()
        for step in range(6):
            x = torch.randint(0, 256, (512, 1, 28, 28), device=dev, generator=g, dtype=torch.uint8)
            y = (x.float().mean((1, 2, 3)) > 127.5).long() + 2 * (x[:, 0, 0, 0] > 127).long()
            eng.run(x, y, True)
        eng.end_pass()
        results[comm] = torch.cat([p.detach().float().reshape(-1).clone() for p in model.parameters()])
        gathered = [torch.empty_like(results[comm]) for _ in range(world)]
        dist.all_gather(gathered, results[comm])
        assert all(torch.equal(gathered[0], t) for t in gathered), f"ranks diverged with comm={comm}"
    diff = (results["pg"] - results["fused"]).abs().max().item()
    scale = results["fused"].abs().max().item()
    assert diff <= 1e-5 * scale, (diff, scale)
    return diff


def test_engine_pg_backend_waits_for_side_stream_gradients():
    _need_gpus(2)
    print("max |pg - fused| parameter difference:", run_gpu_ranks(_engine_pg_matches_fused, 2))


def _buffers_broadcast_by_our_kernel(rank, world):
    import torch.nn as nn
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    dev = torch.device("cuda", rank)
    torch.manual_seed(rank)
    net = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 8, 1), nn.BatchNorm2d(8), nn.ReLU(),
                        nn.Flatten(), nn.Linear(8 * 6 * 6, 4)).to(dev)          # two BatchNorms: their re-homed buffers
    bn = net[1]                                                                    # must not share a version counter
    bn.running_mean.fill_(float(rank + 1)); bn.running_var.fill_(float(10 * (rank + 1))); bn.num_batches_tracked.fill_(rank + 5)
    ddp = DDP(net, device_ids=[rank], comm="fused", bucket_cap_mb=0.001)
    # constructor: rank 0's parameters and buffers everywhere, through our broadcast kernel
    assert float(bn.running_mean[0]) == 1.0 and float(bn.running_var[3]) == 10.0 and int(bn.num_batches_tracked) == 5
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    ref = flat.clone(); dist.broadcast(ref, 0)
    assert torch.equal(flat, ref)
    b0 = ddp.backend.broadcasts
    # every grad-enabled forward re-broadcasts rank 0's buffers (torch DDP semantics, reference utils/process.py:20);
    # eval-mode BatchNorm so the forward itself does not touch the statistics again
    ddp.eval()
    for step in range(3):
        bn.running_mean.add_(rank * 0.5 + step)               # ranks drift ...
        bn.num_batches_tracked.add_(rank)
        x = torch.randn(4, 3, 8, 8, device=dev)
        ddp(x).sum().backward()
        torch.cuda.synchronize()
        got = bn.running_mean.clone(); ref = got.clone(); dist.broadcast(ref, 0)
        assert torch.equal(got, ref), f"step {step}: rank {rank} running_mean differs from rank 0"
        assert int(bn.num_batches_tracked) == 5                # ... and are pulled back to rank 0's values
    assert ddp.backend.broadcasts - b0 == 3
    ddp.train()                                                # train-mode backward through both BatchNorms (saved statistics)
    ddp(torch.randn(4, 3, 8, 8, device=dev)).sum().backward()
    sd = ddp.state_dict()
    assert sd["module.1.num_batches_tracked"].dtype == torch.int64
    return ddp.backend.provider


def test_buffer_and_parameter_broadcast_use_the_symm_kernel():
    _need_gpus(2)
    run_gpu_ranks(_buffers_broadcast_by_our_kernel, min(torch.cuda.device_count(), 4))


def _odd_world_buckets(rank, world):
    """buckets are padded to 32 B x world, so 3 / 5 / 6 / 7 GPUs work with the fused back end."""
    import torch.nn as nn
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(37, 301), nn.ReLU(), nn.Linear(301, 10)).to(dev)
    for wire in ("fp32", "bf16"):
        ddp = DDP(net, device_ids=[rank], comm="fused", bucket_cap_mb=0.005, comm_wire=wire, tail_bucket_kb=1)
        assert all(b.numel % (8 * world) == 0 for b in ddp.buckets) and len(ddp.buckets) >= 2
        x = torch.randn(8, 37, device=dev) + rank
        ddp.zero_grad()
        ddp(x).sum().backward()
        torch.cuda.synchronize()
        for p in net.parameters():
            ref = p.grad.clone(); dist.broadcast(ref, 0)
            assert torch.equal(p.grad, ref)
    return True


def test_fused_backend_on_three_gpus():
    _need_gpus(3)
    run_gpu_ranks(_odd_world_buckets, 3)


def _engine_deterministic_world(rank, world):
    """DDPT_DETERMINISTIC=1 at world >= 2: two runs of the same program end with bit-identical parameters (the fused
    all-reduce sums in a fixed order: rank order in the one-shot kernel, owner order / the switch in the two-shot)."""
    import torch.nn as nn
    os.environ["DDPT_DETERMINISTIC"] = "1"
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    from ddpt_b200.utils import same_seeds
    dev = torch.device("cuda", rank)
    runs = []
    for attempt in range(2):
        amp.reset()
        same_seeds(5)
        model = Toy_Net().to(dev)
        opt = Lookahead(SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True), k=3, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        ddp = DDP(model, device_ids=[rank], comm="fused")
        loss_fn = nn.CrossEntropyLoss()
        assert ops.attach_step_engine(ddp, loss_fn, opt) and model._fused.deterministic
        eng = ddp._ddpt_step_engine
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        ddp.train(); eng.begin_pass()
        for step in range(12):
            x = torch.randint(0, 256, (256, 1, 28, 28), device=dev, generator=g, dtype=torch.uint8)
            y = (x.float().mean((1, 2, 3)) > 127.5).long() + 2 * (x[:, 0, 0, 0] > 127).long()
            eng.run(x, y, True)
        eng.end_pass()
        runs.append(torch.cat([p.detach().float().reshape(-1).clone() for p in model.parameters()]))
    assert torch.equal(runs[0].view(torch.int32), runs[1].view(torch.int32)), float((runs[0] - runs[1]).abs().max())
    return True


def test_deterministic_mode_world2_is_bit_reproducible():
    _need_gpus(2)
    run_gpu_ranks(_engine_deterministic_world, 2)


def _graphed_generic_ddp(rank, world):
    """engine.GraphedTrainStep on a generic autograd model under our DDP: forward, backward with the reducer's fused
    all-reduce launched from autograd hooks, loss scaling and the fused optimizer captured as ONE CUDA graph; ranks
    stay bit-identical and the loss follows the eager trajectory."""
    import torch.nn as nn
    import torch.nn.functional as F
    from ddpt_b200 import amp
    from ddpt_b200.engine import GraphedTrainStep
    from ddpt_b200.ops.conv_modules import TensorCoreConv1x1
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    from ddpt_b200.utils import same_seeds
    dev = torch.device("cuda", rank)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = nn.Sequential(nn.Conv2d(3, 64, 3, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU())
            self.pw = TensorCoreConv1x1(64, 128, 1, bias=False)
            self.bn = nn.BatchNorm2d(128)
            self.head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(128, 10))

        def forward(self, x):
            return self.head(F.relu(self.bn(self.pw(self.stem(x)))))

    g = torch.Generator(device=dev).manual_seed(50 + rank)
    xs = [torch.randn(16, 3, 16, 16, device=dev, generator=g).contiguous(memory_format=torch.channels_last) for _ in range(8)]
    ys = [(x.mean((1, 2, 3)) > 0).long() + 2 * (x[:, 0, 0, 0] > 0).long() for x in xs]
    loss_fn = nn.CrossEntropyLoss()
    out = {}
    for mode in ("graph", "eager"):
        amp.reset()
        same_seeds(9)
        model = Net().to(dev).to(memory_format=torch.channels_last)
        opt = Lookahead(SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True), k=3, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        cap = torch.cuda.Stream()
        with torch.cuda.stream(cap):
            ddp = DDP(model, device_ids=[rank], comm="fused", bucket_cap_mb=0.05)
        torch.cuda.synchronize()
        step = GraphedTrainStep(ddp, loss_fn, opt, xs[0], ys[0], warmup=2, use_graph=(mode == "graph"), stream=cap)
        losses = [float(step(x, y)) for x, y in zip(xs, ys)]
        assert (step.replays == len(xs)) == (mode == "graph")
        flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
        ref = flat.clone(); dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), f"ranks diverged ({mode})"
        out[mode] = losses
    for a, b in zip(out["graph"], out["eager"]):
        assert abs(a - b) <= 0.05 * abs(b) + 5e-3, (out["graph"], out["eager"])
    return out["graph"][-1]


def test_graphed_train_step_with_ddp_hooks_world2():
    _need_gpus(2)
    run_gpu_ranks(_graphed_generic_ddp, 2)


def _bf16_wire_under_graph(rank, world):
    """Regression: the bf16 wire's cast / expand phases must use the same CTA-to-vector mapping as the reduce phase (the
    barriers only pair CTA b with CTA b).  The first version passed every eager test and produced gradients off by
    world x loss_scale under CUDA-graph replay, where co-running kernels delay individual CTAs."""
    import torch.nn as nn
    from ddpt_b200 import amp, ops
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel as DDP
    from ddpt_b200.utils import same_seeds
    dev = torch.device("cuda", rank)
    grads = {}
    for wire in ("fp32", "bf16"):
        amp.reset(); same_seeds(3)
        model = Toy_Net().to(dev)
        opt = Lookahead(SGD(model.parameters(), lr=0.0, momentum=0.9, nesterov=True), k=10, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        ddp = DDP(model, device_ids=[rank], comm="fused", comm_wire=wire)
        assert ops.attach_step_engine(ddp, nn.CrossEntropyLoss(), opt)
        eng = ddp._ddpt_step_engine
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        x = torch.randint(0, 256, (1024, 1, 28, 28), device=dev, generator=g, dtype=torch.uint8)
        y = torch.randint(0, 10, (1024,), device=dev, generator=g)
        ddp.train(); eng.begin_pass()
        for it in range(6):                                   # eager, capture + replay, 4 more replays
            eng.run(x, y, True)
            torch.cuda.synchronize()
            grads[(wire, it)] = ddp.arena.grad_arena.clone()
        assert eng.graph_replays >= 5
    worst = 0.0
    for it in range(6):
        for bk in ddp.buckets:
            a = grads[("fp32", it)][bk.start:bk.start + bk.numel]
            b = grads[("bf16", it)][bk.start:bk.start + bk.numel]
            rel = float((b - a).norm() / a.norm().clamp_min(1e-30))
            assert rel < 2e-2, f"iter {it} bucket {bk.index}: bf16-wire gradients differ from fp32-wire by {rel:.3e}"
            worst = max(worst, rel)
    return worst


def test_bf16_wire_matches_fp32_wire_under_graph_replay():
    _need_gpus(2)
    print("max relative gradient difference bf16 vs fp32 wire:", run_gpu_ranks(_bf16_wire_under_graph, 2))
