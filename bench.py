"""Headline benchmark: MNIST Toy_Net training images/sec, whole job (BASELINE.json).

    python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Config: Toy_Net, per-GPU batch 1024 (reference default), bf16 AMP O1 with dynamic loss scaling,
Lookahead(k=10, alpha=0.5)-wrapped Nesterov SGD, DDP gradient buckets of 25 MB, synthetic
1x28x28 uint8 data, random-init weights.  Weak scaling: per-GPU batch fixed.

Two numbers per run:
  value  device-timed (CUDA events, max over ranks) throughput of K full training steps with the
         batches already resident on the device;
  e2e    K steps through the public API ``iterate_loader(loader, model, loss, rank, opt, training=True)``
         fed by the repo's own input pipeline: ``BatchLoader(mode="pinned", sampler=ShardedSampler)`` over a
         60 000-image synthetic dataset in host memory (per-epoch permutation, native threaded gather into
         pinned buffers, H2D inside the timed region), loss/accuracy copied back every step.
``--impl reference`` runs the unmodified reference sources (baseline/_ref/src) through their own
``iterate_loader`` on the same config (see baseline/README.md for the offline shims).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries exactly one JSON line: NCCL's banner / debug output (it defaults to stdout) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "MNIST Toy_Net training images/sec (whole job, device-timed, max over ranks)"
BATCH = 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no_graph", action="store_true")
    ap.add_argument("--comm", default="fused", choices=("fused", "auto", "pg"),
                    help="gradient all-reduce back end at N>1: fused = our NVLink kernels (hard failure if unavailable)")
    ap.add_argument("--comm_wire", default="fp32", choices=("fp32", "bf16"))
    ap.add_argument("--opt_level", default="O1", choices=("O1", "O1-fp16", "O1-fp8"),
                    help="O1: bf16 (headline).  O1-fp8: conv.7/conv.10/dense.1 forward on block-scaled fp8 tensor cores")
    ap.add_argument("--skip_e2e", action="store_true", help="experiments only: device-timed loop, no end-to-end loops")
    ap.add_argument("--census", default=None, help="write a per-step kernel census (CUPTI) of our arm to this JSON file")
    ap.add_argument("--local_rank", "--local-rank", type=int, default=None)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------------
def dist_env(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", args.local_rank if args.local_rank is not None else 0))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local))
    return rank, local, world


def max_over_ranks(ms: float, world: int) -> float:
    import torch
    import torch.distributed as dist
    if world == 1:
        return ms
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier_sync(world: int) -> None:
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def synthetic_batches(n_batches: int, batch: int, seed: int, pinned: bool):
    """uint8 [B,1,28,28] images + int64 labels, class-dependent so the loss is meaningful."""
    import torch
    g = torch.Generator().manual_seed(seed)
    protos = torch.rand((10, 28, 28), generator=g)
    out = []
    for _ in range(n_batches):
        y = torch.randint(0, 10, (batch,), generator=g)
        x = ((0.6 * protos[y] + 0.4 * torch.rand((batch, 28, 28), generator=g)) * 255).to(torch.uint8)
        x = x.view(batch, 1, 28, 28)
        if pinned:
            x, y = x.pin_memory(), y.pin_memory()
        out.append((x, y))
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.nn as nn
    from ddpt_b200 import amp, ops
    from ddpt_b200.engine import iterate_loader
    from ddpt_b200.models import Toy_Net
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.parallel import DistributedDataParallel
    from ddpt_b200.utils import same_seeds

    rank, local, world = dist_env(args)
    device = torch.device("cuda", local)
    ops.require_native()
    same_seeds(42)
    model = Toy_Net().to(device)
    sgd = SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
    opt = Lookahead(sgd, k=10, alpha=0.5)
    loss_fn = nn.CrossEntropyLoss()
    amp.reset()
    model, opt = amp.initialize(model, optimizers=opt, opt_level=args.opt_level)
    if world > 1:
        model = DistributedDataParallel(model, device_ids=[local], output_device=local, bucket_cap_mb=25.0,
                                        comm=args.comm, comm_wire=args.comm_wire)
    if not ops.attach_step_engine(model, loss_fn, opt, use_graph=not args.no_graph):
        raise RuntimeError("step engine could not be attached (native path missing)")
    engine = model._ddpt_step_engine
    B, K, W = args.batch, args.steps, args.warmup
    model.train()

    # ---- (1) device-resident batches: kernel-only throughput ----------------------------------------
    nres = 8                                                   # 8 x 0.8 MB inputs cycle; activations (>1 GB/step) exceed L2
    dev_batches = [(x.to(device), y.to(device)) for x, y in synthetic_batches(nres, B, 100 + rank, False)]
    engine.begin_pass()
    for i in range(max(W, 3)):
        engine.run(*dev_batches[i % nres], True)
    barrier_sync(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    replays0 = engine.graph_replays
    start.record()
    for i in range(K):
        engine.run(*dev_batches[i % nres], True)
    stop.record()
    barrier_sync(world)
    ms_dev = max_over_ranks(start.elapsed_time(stop), world)
    clocks = sampler.stop() if rank == 0 else {}
    replays = engine.graph_replays - replays0
    launches_per_step = engine.fused_model.count_launches(B) + (len(model.buckets) if world > 1 else 0)

    # ---- (2) end to end through the public API and the repo's own input pipeline ------------------------------
    # A synthetic 60 000-image MNIST-shaped dataset lives in host memory.  Every epoch the DistributedSampler-parity
    # ShardedSampler draws this rank's permutation, the native loader (csrc/data/host_loader.cpp) gathers each batch
    # row by row into pinned buffers on worker threads, iterate_loader() copies it to the device (H2D on a copy
    # stream) and runs the step; loss/accuracy accumulators are copied back every step (async D2H) and read at the
    # end of every epoch (one host sync per iterate_loader call, like the reference's epoch loop, utils/train.py:77).
    # The timed region covers exactly K optimizer steps, epoch boundaries and tail batches (drop_last=False) included
    # -- the same work the reference arm's DataLoader(DistributedSampler) + iterate_loader does.
    from ddpt_b200.data import BatchLoader, ShardedSampler, synthetic_mnist
    engine.host_metrics_every_step = True
    train_set = synthetic_mnist(60000, seed=1)
    sampler = ShardedSampler(len(train_set), num_replicas=world, rank=rank)

    def run_steps(loader, n_steps, first_epoch):
        """n_steps optimizer steps through iterate_loader, epoch after epoch; returns (images, epochs, last metrics)."""
        left, epoch, images, out = n_steps, first_epoch, 0, None
        while left > 0:
            sampler.set_epoch(epoch)
            loader.max_batches = left
            before = loader.samples_yielded
            out = iterate_loader(loader, model, loss_fn, local, opt, training=True)
            left -= len(loader)
            images += loader.samples_yielded - before
            epoch += 1
        return images, epoch, out

    results = {}
    for mode in (() if args.skip_e2e else ("pinned", "device")):
        loader = BatchLoader(train_set, B, sampler, device, mode=mode)
        loader.quiet = True
        loader.max_batches = None
        per_epoch = len(loader)
        # warm-up: two whole epochs so the tail-batch size has its CUDA graph too (eager, capture), then W steps
        _, ep, _ = run_steps(loader, 2 * per_epoch + max(W, 3), 0)
        barrier_sync(world)
        h2d0 = loader.h2d_bytes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        images, ep_end, (loss, acc, lr) = run_steps(loader, K, ep)
        e1.record()
        barrier_sync(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world)
        tot = torch.tensor([float(images)], device=device, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tot)
        results[mode] = {"ms": ms, "images": float(tot.item()), "epochs": ep_end - ep, "loss": loss, "acc": acc,
                         "h2d_per_step": (loader.h2d_bytes - h2d0) / K}
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": world * B * K / ms_dev * 1e3, "n_gpus": world, "steps": K,
                              "ms_per_step": ms_dev / K, "e2e": None, "clocks": clocks, "experiment": True,
                              "config": {"comm": getattr(model, "comm", "none"),
                                         "comm_wire": getattr(getattr(model, "backend", None), "wire_name", None)}}))
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    ms_e2e = results["pinned"]["ms"]
    loss, acc = results["pinned"]["loss"], results["pinned"]["acc"]
    h2d = int(results["pinned"]["h2d_per_step"])
    d2h = 8

    census = None
    if args.census:
        # kernel census of two graph replays (CUPTI through torch.profiler), outside every timed region
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                engine.run(*dev_batches[i % nres], True)
            torch.cuda.synchronize()
        census = {}
        for ev in prof.events():
            if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.name and "Memcpy" not in ev.name \
                    and "Memset" not in ev.name:
                census[ev.name] = census.get(ev.name, 0) + 1
        census = {k: v / 2 for k, v in sorted(census.items(), key=lambda kv: -kv[1])}
        if rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(args.census)), exist_ok=True)
            json.dump({"n_gpus": world, "comm": getattr(model, "comm", "none"),
                       "comm_wire": getattr(getattr(model, "backend", None), "wire_name", None),
                       "buckets": [b.nbytes for b in getattr(model, "buckets", [])],
                       "kernels_per_step": census}, open(args.census, "w"), indent=1)

    if rank == 0:
        value = world * B * K / ms_dev * 1e3
        e2e = results["pinned"]["images"] / ms_e2e * 1e3
        e2e_device_loader = results["device"]["images"] / results["device"]["ms"] * 1e3
        out = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"O1": "bf16", "O1-fp16": "fp16"}.get(args.opt_level, "bf16 + mxfp8 forward GEMMs"),
            "data": "synthetic uint8 1x28x28, random-init weights", "impl": "ours",
            "config": {"model": "Toy_Net", "global_batch": world * B, "per_gpu_batch": B, "seq_len": None,
                       "parallelism": f"dp{world}", "amp": args.opt_level + " + dynamic loss scale",
                       "optimizer": "Lookahead(k=10,a=0.5)+SGD(nesterov,0.9)", "bucket_cap_mb": 25,
                       "comm": getattr(model, "comm", "none"),
                       "comm_wire": getattr(getattr(model, "backend", None), "wire_name", None),
                       "grad_buckets_bytes": [b.nbytes for b in getattr(model, "buckets", [])],
                       "cuda_graph": not args.no_graph,
                       "l2": "per-step working set > 1 GB of activations (>> 126 MB L2); 8 input batches cycled"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "final_loss": loss, "final_acc": acc,
                    "images": results["pinned"]["images"], "epochs_touched": results["pinned"]["epochs"],
                    "loader": "BatchLoader(mode='pinned', sampler=ShardedSampler) over a 60 000-image synthetic set",
                    "device_resident_loader_value": e2e_device_loader,
                    "note": "K steps through iterate_loader(): per-epoch DistributedSampler-parity permutation, native "
                            "threaded row gather into pinned buffers, H2D on a copy stream, per-step async D2H of "
                            "loss/acc, one host sync per epoch; tail batches included (drop_last=False). "
                            "device_resident_loader_value = the same loop with the uint8 dataset resident on the GPU "
                            "(BatchLoader mode='device', reported separately: no per-step H2D)"},
            "gpu_launches": int(launches_per_step * K),
            "graph_replays": int(replays),
            "published_baseline_note": "reference publishes no throughput; ~17.5k img/s derived from its logs on 2x RTX 2080",
        }
        out["vs_baseline"] = value / 17500.0
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm: unmodified sources from baseline/_ref/src
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    src = os.path.join(ROOT, "baseline", "_ref", "src")
    if not os.path.isfile(os.path.join(src, "utils", "process.py")):
        print(json.dumps({"impl": "reference", "unavailable":
                          "baseline/_ref/src missing (run baseline/install_ref.sh; pip cannot install the reference: no setup.py)"}))
        return
    try:
        import torch
        import torch.distributed as dist
        import torch.nn as nn
        sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))      # apex stand-in (not installable offline)
        sys.path.insert(0, src)
        for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
            del sys.modules[name]
        import torchvision  # noqa: F401  (utils.train imports it)
        import utils as ref_utils                                         # the reference's public API
        from utils.process import iterate_loader as ref_iterate
        from torch.nn.parallel import DistributedDataParallel as DDP
        from torch.utils.data import DataLoader, Dataset
        from torch.utils.data.distributed import DistributedSampler
        from apex import amp as ref_amp

        rank, local, world = dist_env(args)
        if world == 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", init_method="env://", rank=0, world_size=1,
                                    device_id=torch.device("cuda", local))
        B, K, W = args.batch, args.steps, args.warmup

        class SyntheticMNIST(Dataset):                                     # ToTensor-equivalent samples
            """``n`` samples served from a pool of 8 batches (the index stream still comes from the reference's
            DistributedSampler; a materialised 1.6 M-image tensor per rank would only measure the host's RAM)."""

            def __init__(self, n, seed):
                xs, ys = zip(*synthetic_batches(8, B, seed, False))
                self.x = torch.cat(xs).float().div_(255.0)
                self.y = torch.cat(ys)
                self.n = n

            def __len__(self):
                return self.n

            def __getitem__(self, i):
                j = i % self.x.shape[0]
                return self.x[j], int(self.y[j])

        def loader(n_steps, seed):
            ds = SyntheticMNIST(n_steps * B * world, seed)
            sampler = DistributedSampler(ds)
            return DataLoader(ds, batch_size=B, shuffle=False, pin_memory=True, sampler=sampler)

        ref_utils.same_seeds(42)
        model = ref_utils.Toy_Net().to(local)
        optimizer = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
        lookahead = ref_utils.Lookahead(optimizer=optimizer, k=10, alpha=0.5)
        # torch>=2 compat for the vendored Lookahead (never calls Optimizer.__init__): give it the
        # attributes Optimizer.zero_grad needs -- done from outside, the source file is untouched.
        from collections import OrderedDict
        lookahead.defaults = {}
        lookahead._zero_grad_profile_name = "Optimizer.zero_grad#Lookahead.zero_grad"
        for hook_attr in ("_optimizer_step_pre_hooks", "_optimizer_step_post_hooks",
                          "_optimizer_state_dict_pre_hooks", "_optimizer_state_dict_post_hooks",
                          "_optimizer_load_state_dict_pre_hooks", "_optimizer_load_state_dict_post_hooks"):
            setattr(lookahead, hook_attr, OrderedDict())
        loss_function = nn.CrossEntropyLoss()
        model, apex_optimizer = ref_amp.initialize(model, optimizers=lookahead, opt_level="O1")
        parallel_model = DDP(model, device_ids=[local], output_device=local)
        parallel_model.train()

        import tqdm as _tqdm                                               # keep bars off the log
        import utils.process as ref_process
        ref_process.tqdm = lambda it, **kw: it

        ref_iterate(loader(max(W, 3), 300 + rank), parallel_model, loss_function, local, apex_optimizer, True)
        timed = loader(K, 400 + rank)
        barrier_sync(world)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, acc, lr = ref_iterate(timed, parallel_model, loss_function, local, apex_optimizer, True)
        e1.record()
        barrier_sync(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world)
        clocks = sampler.stop() if rank == 0 else {}
        if rank == 0:
            value = world * B * K / ms * 1e3
            print(json.dumps({
                "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
                "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": value / 17500.0, "dtype": os.environ.get("REF_AMP_DTYPE", "fp16"),
                "data": "synthetic 1x28x28, random-init weights", "impl": "reference",
                "config": {"model": "Toy_Net", "global_batch": world * B, "per_gpu_batch": B, "seq_len": None,
                           "parallelism": f"dp{world}", "amp": "apex-O1 shim over torch.autocast + dynamic loss scale",
                           "path": "unmodified reference utils.process.iterate_loader + torch DDP/NCCL + DataLoader"},
                "clocks": clocks,
                "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": B * 784 * 4 + B * 8,
                        "d2h_bytes_per_step": 8, "note": "the reference loop is end to end by construction"},
                "gpu_launches": 0, "final_loss": loss}))
        dist.destroy_process_group()
    except Exception as exc:
        import traceback
        traceback.print_exc(file=sys.stderr)
        print(json.dumps({"impl": "reference", "unavailable": f"{type(exc).__name__}: {exc}"[:300]}))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
