"""Entry point (reference: run.py).

    python -m torch.distributed.launch --nproc_per_node=2 --master_port=6666 run.py -n=DDP_warmup
    torchrun --nproc-per-node 8 run.py -n DDP_warmup
    python run.py -n smoke --synthetic -e 2          # single process, GPU or CPU

One process per GPU; NCCL/NVLink when CUDA is present, gloo otherwise.
"""
from ddpt_b200.config import parse_args
from ddpt_b200.parallel import init_distributed, shutdown_distributed


def main() -> None:
    args = parse_args()
    print(f"Running DDP on rank: {args.local_rank}")
    init_distributed(args.local_rank, backend=args.backend)
    try:
        from ddpt_b200.engine import train
        train(args)
    finally:
        shutdown_distributed()


if __name__ == "__main__":
    main()
