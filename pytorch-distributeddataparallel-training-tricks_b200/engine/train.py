"""Training driver (reference: utils/train.py:23-118).

Order of operations, callbacks and their hyper-parameters follow the reference's
*code* (not its README, quirk Q9): Nesterov SGD(momentum 0.9) wrapped in
Lookahead(k=10, alpha=0.5); warm-up LambdaLR (linear | cosine) constructed before
training so epoch 0 runs at lr 0 (Q3) and stepped only while ``epoch <=
warmup_epochs`` (Q2); ReduceLROnPlateau(min, 0.1, patience 6) on the validation
loss every epoch; EarlyStopping(patience 30) whose ``break`` precedes the
checkpoint block (Q6); rank-0 TensorBoard scalars; best-validation-accuracy
weights-only checkpoint ``{model_path}/{exp_name}.pt``; validation unsharded on
every rank through the DDP wrapper (Q4); rank-local train metrics (Q5).

Extensions, all opt-in or invisible: CPU/gloo execution, synthetic data,
``--resume`` full-state checkpoints, rank-0-agreed early stopping, sharded
validation, quiet logging, per-epoch step timing.
"""
from __future__ import annotations

import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import amp
from ..data import make_loaders
from ..models import Toy_Net, resnet50
from ..optim import SGD, LambdaLR, Lookahead, ReduceLROnPlateau, warmup_lambda
from ..parallel import DistributedDataParallel
from ..utils import CudaStepTimer, EarlyStopping, ScalarLogger, nvtx_range, same_seeds, save_weights
from ..utils.checkpoint import load_resume_state, save_resume_state
from .process import iterate_loader


def _runtime(args):
    """(device, rank, world) -- works with or without an initialised process group."""
    live = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if live else 0
    world = dist.get_world_size() if live else 1
    if torch.cuda.is_available():
        device = torch.device("cuda", args.local_rank if args.local_rank is not None else 0)
    else:
        device = torch.device("cpu")
    return device, rank, world, live


def _allreduce_mean(values, device):
    t = torch.tensor(values, dtype=torch.float64, device=device if device.type == "cuda" else "cpu")
    dist.all_reduce(t)
    return (t / dist.get_world_size()).tolist()


def train(args) -> dict:
    device, rank, world, live = _runtime(args)
    is_main = rank == 0
    quiet = getattr(args, "quiet", False)

    # ---- data (train sharded by rank, valid redundant) -----------------------------
    same_seeds(args.seed_num, deterministic=True if getattr(args, "deterministic", False) else None)
    train_loader, valid_loader, train_sampler = make_loaders(args, device, world, rank)
    train_loader.quiet = valid_loader.quiet = quiet or not is_main and getattr(args, "quiet", False)

    print(f"Now Training: {args.exp_name}")
    # ---- model ------------------------------------------------------------------------
    same_seeds(args.seed_num)
    arch = getattr(args, "arch", "toy_net")
    model = Toy_Net() if arch == "toy_net" else resnet50(10)
    model = model.to(device)

    os.makedirs(os.path.join(args.model_path, "logs"), exist_ok=True)
    latest_model_path = os.path.join(args.model_path, args.exp_name)
    optimizer = SGD(model.parameters(), lr=args.learning_rate, momentum=0.9, nesterov=True)
    lookahead = Lookahead(optimizer=optimizer, k=10, alpha=0.5)
    loss_function = nn.CrossEntropyLoss()
    best_valid_acc = 0.0

    # ---- callbacks -----------------------------------------------------------------------
    scheduler_wu = LambdaLR(optimizer, warmup_lambda(args.warmup_type, args.warmup_epochs, args.epochs))
    scheduler_re = ReduceLROnPlateau(optimizer, mode="min", factor=0.1, patience=6, verbose=True)
    early_stopping = EarlyStopping(patience=30, verbose=True)

    # ---- mixed precision, then data parallel (order matters, as in the reference) --------
    amp.reset()
    model, apex_optimizer = amp.initialize(model, optimizers=lookahead,
                                           opt_level=getattr(args, "opt_level", "O1"),
                                           fused=not getattr(args, "no_fused", False))
    if live:
        parallel_model = DistributedDataParallel(
            model, device_ids=[args.local_rank] if device.type == "cuda" else None,
            output_device=args.local_rank if device.type == "cuda" else None,
            bucket_cap_mb=getattr(args, "bucket_cap_mb", 25.0), comm=getattr(args, "comm", "auto"),
            comm_wire=getattr(args, "comm_wire", "fp32"))
    else:
        parallel_model = model
    if device.type == "cuda":
        from .. import ops
        ops.attach_step_engine(parallel_model, loss_function, apex_optimizer,
                               use_graph=not getattr(args, "no_graph", False))

    start_epoch = 0
    resume_path = getattr(args, "resume", None)
    schedulers = {"warmup": scheduler_wu, "plateau": scheduler_re}
    if resume_path:
        info = load_resume_state(resume_path, model=parallel_model, optimizer=apex_optimizer,
                                 schedulers=schedulers, early_stopping=early_stopping,
                                 scaler=amp.current_state().scaler)
        start_epoch, best_valid_acc = info["epoch"] + 1, info["best_valid_acc"]

    tb = ScalarLogger(os.path.join(args.model_path, "logs", args.exp_name)) if is_main else None
    history = []
    profile = getattr(args, "profile", False)
    train_timer, valid_timer = CudaStepTimer(), CudaStepTimer()      # device-side epoch timing (--profile)
    try:
        for epoch in range(start_epoch, args.epochs):
            t0 = time.time()
            train_sampler.set_epoch(epoch)

            parallel_model.train()
            with nvtx_range(f"train/epoch{epoch}", profile), train_timer.region():
                train_loss, train_acc, curr_lr = iterate_loader(
                    loader=train_loader, model=parallel_model, loss_function=loss_function,
                    local_rank=args.local_rank, apex_optimizer=apex_optimizer, training=True)

            parallel_model.eval()
            with torch.no_grad(), nvtx_range(f"valid/epoch{epoch}", profile), valid_timer.region():
                valid_loss, valid_acc = iterate_loader(
                    loader=valid_loader, model=parallel_model, loss_function=loss_function,
                    local_rank=args.local_rank, apex_optimizer=None, training=False)

            if live and world > 1:
                if getattr(args, "global_metrics", False):
                    train_loss, train_acc = _allreduce_mean([train_loss, train_acc], device)
                if getattr(args, "shard_valid", False):
                    valid_loss, valid_acc = _allreduce_mean([valid_loss, valid_acc], device)

            if tb is not None:
                tb.epoch(epoch, curr_lr, train_loss, valid_loss, train_acc, valid_acc)
            if is_main or not quiet:
                print(f"epoch: {epoch:03d}/{args.epochs}, time: {time.time() - t0:.2f}s, "
                      f"learning_rate: {curr_lr}, train_loss: {train_loss:.4f}, train_acc: {train_acc:.4f}, "
                      f"valid_loss: {valid_loss:.4f}, valid_acc: {valid_acc:.4f}")
            history.append(dict(epoch=epoch, lr=curr_lr, train_loss=train_loss, train_acc=train_acc,
                                valid_loss=valid_loss, valid_acc=valid_acc, seconds=time.time() - t0))

            if epoch <= args.warmup_epochs or getattr(args, "fix_cosine", False):
                scheduler_wu.step()
            scheduler_re.step(valid_loss)
            early_stopping(valid_loss)
            stop = early_stopping.agree() if live and world > 1 else early_stopping.early_stop
            if stop:
                break

            if is_main and valid_acc > best_valid_acc:
                best_valid_acc = valid_acc
                save_weights(parallel_model, f"{latest_model_path}.pt")
            if is_main and getattr(args, "save_resume", False):
                save_resume_state(f"{latest_model_path}.resume.pt", model=parallel_model,
                                  optimizer=apex_optimizer, schedulers=schedulers,
                                  early_stopping=early_stopping, scaler=amp.current_state().scaler,
                                  epoch=epoch, best_valid_acc=best_valid_acc)
    finally:
        if tb is not None:
            tb.close()
    if profile and is_main and device.type == "cuda":
        steps = len(train_loader)
        t = train_timer.summary()
        if t.get("n"):
            print(f"[profile] train pass: median {t['median_ms']:.2f} ms device time = "
                  f"{t['median_ms'] / max(steps, 1):.3f} ms/step, "
                  f"{steps * args.batch_size / t['median_ms'] * 1e3:.0f} images/s/GPU; "
                  f"valid pass median {valid_timer.summary().get('median_ms', float('nan')):.2f} ms")
    return {"history": history, "best_valid_acc": best_valid_acc, "model": parallel_model,
            "optimizer": apex_optimizer}
