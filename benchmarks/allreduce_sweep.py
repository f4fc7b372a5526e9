"""Gradient all-reduce bandwidth sweep (BASELINE.json config 5): fused NVLink kernel vs NCCL.

    torchrun --nproc-per-node N benchmarks/allreduce_sweep.py [--out profiles/allreduce_sweep_nN.json]

Buckets 64 KB - 256 MB of fp32 gradients.  Device-timed (CUDA events), max over ranks, L2 flushed between
iterations.  Every fused number includes the 1/(world x loss_scale) multiply and the overflow scan; the bf16 rows
also include the fp32 -> bf16 cast on send and the expansion back into the fp32 arena (the bucket is the same
number of gradients, the wire carries half the bytes).  bus GB/s = 2 (n-1)/n x fp32 bytes / time (the NCCL
convention, so bf16-wire rows can exceed the link rate); the roofline is the measured 770 GB/s per direction per
GPU (900 nominal)."""
import argparse, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--max_mb", type=int, default=256)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from ddpt_b200.ops import require_native
    from ddpt_b200.parallel.symm import SymmetricRegion
    ext = require_native()
    max_elems = a.max_mb * 1024 * 1024 // 4
    region = SymmetricRegion(max_elems, dev)
    state = torch.tensor([1.0, 1.0, 0.0, 0.0, 0.0], device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    nccl_buf = torch.randn(max_elems, device=dev)
    nccl_bf16 = nccl_buf.bfloat16()
    region.tensor.copy_(nccl_buf)

    def timed(fn, iters, reps=1):
        """Median device time of one call.  `reps` back-to-back calls between the events take the host's launch
        latency (~10 us of Python per call) out of small-bucket numbers: what a CUDA graph replays is the kernel."""
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            torch.cuda.synchronize(); dist.barrier()
            s, e = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record()
            for _ in range(reps):
                fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / reps)
        t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def fused(n, algo, wire, blocks):
        return lambda: ext.fused_allreduce(region.peer_ptrs, region.signal_ptrs, region.mc_ptr, region.scratch_ptrs,
                                           region.mc_scratch_ptr, rank, 0, n, state, state[2:3], 0, algo, wire, blocks,
                                           20000)

    rows = []
    size = 64 * 1024
    while size <= a.max_mb * 1024 * 1024:
        n = size // 4
        row = {"bytes": size}
        reps = 10 if size <= 4 * 1024 * 1024 else 1       # small buckets: pipelined launches (kernel time, not host time)
        variants = {"f32_auto": (0, 0, 0), "bf16_auto": (0, 1, 0), "f32_twoshot_p2p": (2, 0, 0),
                    "bf16_twoshot_p2p": (2, 1, 0)}
        if region.mc_ptr and n > 64 * 1024:
            variants["f32_multimem"] = (3, 0, 0)
            variants["bf16_multimem"] = (3, 1, 0)
        if size >= 4 * 1024 * 1024:
            variants["f32_auto_b128"] = (0, 0, 128)
            variants["bf16_auto_b128"] = (0, 1, 128)
            variants["f32_auto_b32"] = (0, 0, 32)
        for name, (algo, wire, blocks) in variants.items():
            ms = timed(fused(n, algo, wire, blocks), a.iters, reps)
            row[f"fused_{name}_us"] = round(ms * 1e3, 2)
            row[f"fused_{name}_busGBs"] = round(2 * (world - 1) / world * size / (ms * 1e-3) / 1e9, 1)
        view = nccl_buf[:n]
        ms = timed(lambda: dist.all_reduce(view), a.iters, reps)
        row["nccl_us"] = round(ms * 1e3, 2)
        row["nccl_busGBs"] = round(2 * (world - 1) / world * size / (ms * 1e-3) / 1e9, 1)
        v16 = nccl_bf16[:n]
        ms = timed(lambda: dist.all_reduce(v16), a.iters, reps)
        row["nccl_bf16_us"] = round(ms * 1e3, 2)          # same gradient count sent as bf16 (no cast / scale / check)
        rows.append(row)
        size *= 4
    if rank == 0:
        out = {"world": world, "provider": region.provider, "multicast": bool(region.mc_ptr), "rows": rows,
               "link_GBs_measured_per_direction": 770, "link_GBs_nominal": 900,
               "note": "auto = one-shot push <= 256 KiB, else two-shot: peer loads at 2 GPUs, multimem (NVLS) from 4; "
                       "buckets <= 4 MB are timed as 10 back-to-back launches / 10 (host launch latency excluded, as in "
                       "a CUDA graph); _bN = N CTAs of 256 threads (default 64)"}
        print(json.dumps(out, indent=1))
        if a.out:
            os.makedirs(os.path.dirname(a.out), exist_ok=True)
            json.dump(out, open(a.out, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
