"""Command-line surface (reference: run.py:7-20).

Flag names, short options, defaults and types are identical to the reference so
existing launch lines keep working.  Three launcher spellings for the local
rank are accepted (``--local_rank`` from torch<=1.x launchers, ``--local-rank``
from torch>=2.0, and the ``LOCAL_RANK`` environment variable exported by
``torchrun``); the reference only knew the first and died with argparse exit 2
under a modern launcher (shim 2).

Framework-only knobs live behind long options that the reference never used, so
they cannot collide.
"""
from __future__ import annotations

import argparse
import os
from typing import Optional, Sequence

# (short, long, default, type, help) -- the ten reference flags, in order.
_REFERENCE_FLAGS = (
    ("-n", "--exp_name", "DDP_warmup", str, "name of experiment"),
    ("-l", "--learning_rate", 1e-1, float, "learning rate"),
    ("-b", "--batch_size", 1024, int, "batch size (per GPU)"),
    ("-e", "--epochs", 500, int, "epochs"),
    ("-w", "--warmup_epochs", 10, int, "epochs for warmup"),
    ("-t", "--warmup_type", "linear", str, "warmup type: linear | cosine"),
    ("-s", "--seed_num", 42, int, "number of random seed"),
    ("-d", "--data_path", "./datasets/", str, "path of dataset"),
    ("-p", "--model_path", "./experiment_model/", str, "path of model"),
)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(
        description="B200-native DDP training harness (Toy_Net / MNIST)")
    for short, long_, default, typ, text in _REFERENCE_FLAGS:
        parser.add_argument(short, long_, default=default, type=typ, help=text)
    parser.add_argument("--local_rank", "--local-rank", dest="local_rank", type=int,
                        default=None, help="local rank for DistributedDataParallel")

    ext = parser.add_argument_group("framework extensions (not in the reference)")
    ext.add_argument("--backend", default="auto", choices=("auto", "nccl", "gloo"),
                     help="process-group backend used for rendezvous/plumbing")
    ext.add_argument("--opt_level", default="O1",
                     choices=("O0", "O1", "O1-fp16", "O1-fp8"),
                     help="mixed precision level: O0 fp32, O1 bf16, O1-fp16, O1-fp8 (block scaled)")
    ext.add_argument("--comm", default="auto", choices=("auto", "fused", "pg"),
                     help="gradient all-reduce: fused NVLink kernel or process-group collective")
    ext.add_argument("--comm_wire", default="fp32", choices=("fp32", "bf16"),
                     help="wire format of the fused all-reduce: fp32, or bf16 (cast fused into the send, fp32 accumulate)")
    ext.add_argument("--bucket_cap_mb", default=25.0, type=float, help="gradient bucket size")
    ext.add_argument("--arch", default="toy_net", choices=("toy_net", "resnet50"),
                     help="model family")
    ext.add_argument("--synthetic", action="store_true",
                     help="use synthetic data of MNIST shape instead of reading the dataset")
    ext.add_argument("--loader", default="auto", choices=("auto", "device", "pinned", "torch"),
                     help="input pipeline: device-resident gather, native pinned-host loader, or torch DataLoader")
    ext.add_argument("--shard_valid", action="store_true",
                     help="shard validation across ranks and all-reduce metrics (reference: redundant)")
    ext.add_argument("--global_metrics", action="store_true",
                     help="all-reduce train metrics (reference: rank-local)")
    ext.add_argument("--resume", default=None, type=str,
                     help="resume from a full-state sidecar checkpoint (*.resume.pt)")
    ext.add_argument("--max_steps", default=None, type=int,
                     help="stop every epoch after this many steps (smoke tests)")
    ext.add_argument("--no_fused", action="store_true",
                     help="do not attach the sm_100a fused model/optimizer (library kernels only)")
    ext.add_argument("--fix_cosine", action="store_true",
                     help="keep stepping the warm-up scheduler after warm-up so the cosine decay runs (reference quirk Q2 off)")
    ext.add_argument("--save_resume", action="store_true",
                     help="write a full-state {exp_name}.resume.pt every epoch")
    ext.add_argument("--deterministic", action="store_true",
                     help="bit-reproducible runs: ordered reductions instead of float atomics in the fused kernels")
    ext.add_argument("--no_graph", action="store_true", help="disable CUDA-graph capture of the step")
    ext.add_argument("--profile", action="store_true", help="emit NVTX ranges and step timing summary")
    ext.add_argument("--quiet", action="store_true", help="only rank 0 prints epoch lines; no progress bars")
    return parser


def resolve_local_rank(value: Optional[int]) -> int:
    """Local rank from the flag, else ``LOCAL_RANK``, else 0 (single process)."""
    if value is not None:
        return int(value)
    return int(os.environ.get("LOCAL_RANK", "0"))


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    args = build_parser().parse_args(argv)
    args.local_rank = resolve_local_rank(args.local_rank)
    if args.warmup_type not in ("linear", "cosine"):
        raise SystemExit(f"unknown --warmup_type {args.warmup_type!r} (linear | cosine)")
    return args
