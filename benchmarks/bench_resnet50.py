"""ResNet-50 synthetic 3x224x224 DDP bf16 (BASELINE.json config 4): stresses all-reduce overlap.

    torchrun --nproc-per-node N benchmarks/bench_resnet50.py [--impl ours|torch] [--batch 64]

`ours`  : our DistributedDataParallel (flat arenas, native C++ reducer, fused NVLink all-reduce with the
          amp unscale folded in) + fused Lookahead-SGD + device-resident loss scaler; the 36 pointwise convolutions
          (~52 % of the multiply-accumulates) on our tcgen05 GEMM kernels (forward, dgrad, wgrad), 7x7 / 3x3 on cuDNN;
          the whole step -- forward, backward, bucketed all-reduce, optimizer -- replayed as ONE CUDA graph
          (engine.GraphedTrainStep).  --no_graph / --no_tc switch the two off for ablation.
`torch` : torch DDP + NCCL + torch.optim.SGD + GradScaler-free bf16 autocast (the library baseline)."""
import argparse, json, os, sys
import torch, torch.nn as nn, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no_graph", action="store_true")
    ap.add_argument("--no_tc", action="store_true", help="pointwise convolutions through cuDNN as well")
    ap.add_argument("--comm_wire", default="fp32")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from ddpt_b200.models import resnet50
    torch.manual_seed(0)
    model = resnet50(1000).to(dev).to(memory_format=torch.channels_last)
    x = torch.randn(a.batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch,), device=dev)
    loss_fn = nn.CrossEntropyLoss()
    if a.impl == "ours":
        from ddpt_b200 import amp
        from ddpt_b200.optim import SGD, Lookahead
        from ddpt_b200.parallel import DistributedDataParallel as DDP
        amp.reset()
        opt = Lookahead(SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True), k=10, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        if a.no_tc:
            from ddpt_b200.ops import conv_modules
            for m in model.modules():
                if isinstance(m, conv_modules.TensorCoreConv1x1):
                    m.__class__ = nn.Conv2d
        from ddpt_b200.engine import GraphedTrainStep
        cap = torch.cuda.Stream()                 # DDP is built under the stream the step is captured on (see graphed.py)
        with torch.cuda.stream(cap):
            net = DDP(model, device_ids=[local], bucket_cap_mb=25.0, comm="fused", comm_wire=a.comm_wire) if world > 1 else model
        torch.cuda.synchronize()
        graphed = GraphedTrainStep(net, loss_fn, opt, x, y, warmup=3, use_graph=not a.no_graph, stream=cap)

        def step():
            graphed(x, y)
        comm = getattr(net, "comm", "none")
    else:
        from torch.nn.parallel import DistributedDataParallel as DDP
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
        net = DDP(model, device_ids=[local], bucket_cap_mb=25)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x)
            loss_fn(out.float(), y).backward()
            opt.step()
        comm = "nccl"
    extra = {}
    if a.impl == "ours":
        from ddpt_b200.ops import conv_modules
        extra["share"] = 0.0 if a.no_tc else round(conv_modules.pointwise_flop_share(model), 4)
        extra["calls"] = 0 if a.no_tc else sum(isinstance(m, conv_modules.TensorCoreConv1x1) for m in model.modules())
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(a.steps):
        step()
    e.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) / a.steps
        print(json.dumps({"metric": "ResNet-50 training images/sec (synthetic 3x224x224, bf16)", "impl": a.impl,
                          "value": world * a.batch / ms * 1e3, "unit": "images/s", "n_gpus": world,
                          "ms_per_step": ms, "per_gpu_batch": a.batch, "comm": comm, "cuda_graph": a.impl == "ours" and not a.no_graph,
                          "own_kernel_mac_share": extra.get("share"), "tensor_core_conv_calls_per_step": extra.get("calls"),
                          "grad_bytes": sum(p.numel() for p in model.parameters()) * 4}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
