"""Diagnostic (2 GPUs): gradient arena after one engine step with the fp32 vs the bf16 wire format."""
import os, sys
import torch, torch.nn as nn, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from ddpt_b200 import amp, ops
from ddpt_b200.models import Toy_Net
from ddpt_b200.optim import SGD, Lookahead
from ddpt_b200.parallel import DistributedDataParallel as DDP
from ddpt_b200.utils import same_seeds
res = {}
for wire in ("fp32", "bf16"):
    amp.reset(); same_seeds(3)
    model = Toy_Net().to(dev)
    opt = Lookahead(SGD(model.parameters(), lr=0.0, momentum=0.9, nesterov=True), k=10, alpha=0.5)
    model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
    ddp = DDP(model, device_ids=[local], comm="fused", comm_wire=wire)
    assert ops.attach_step_engine(ddp, nn.CrossEntropyLoss(), opt)
    eng = ddp._ddpt_step_engine
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    x = torch.randint(0, 256, (1024, 1, 28, 28), device=dev, generator=g, dtype=torch.uint8)
    y = torch.randint(0, 10, (1024,), device=dev, generator=g)
    ddp.train(); eng.begin_pass()
    for it in range(3):
        eng.run(x, y, True)
        torch.cuda.synchronize()
        res[(wire, it)] = ddp.arena.grad_arena.clone()
    print(rank, wire, "scaler", amp.current_state().scaler.state.tolist(), flush=True)
for it in range(3):
    a, b = res[("fp32", it)], res[("bf16", it)]
    for bk in ddp.buckets:
        sa, sb = a[bk.start:bk.start + bk.numel], b[bk.start:bk.start + bk.numel]
        ratio = float(sb.norm() / sa.norm().clamp_min(1e-30))
        print(f"rank {rank} iter {it} bucket {bk.index} numel {bk.numel}: |fp32| {float(sa.norm()):.4e} |bf16| {float(sb.norm()):.4e} ratio {ratio:.4f} relerr {float((sb - sa).norm() / sa.norm().clamp_min(1e-30)):.4e}", flush=True)
dist.destroy_process_group()
