"""Gradient bucket planning (what DDP's ``_compute_bucket_assignment_by_size`` +
the one-time rebuild converge to).

Gradients become ready in roughly reverse definition order, so buckets are
planned over the *reversed* parameter list with a small first bucket (1 MiB) so
communication starts early, then ``bucket_cap_mb`` buckets.  For Toy_Net this
yields the reference's steady-state layout -- {dense.4, dense.2, dense.1} =
16 803 880 B, then {all conv/BN2d} = 6 206 976 B -- with no rebuild iteration
and no bucket-index broadcast (N6): the plan is a pure function of the
parameter shapes, hence identical on every rank.

One deliberate difference (B200-first): the **tail split**.  The gradients that
become ready last (the first layers) cannot overlap with any remaining backward
work, so their all-reduce is pure exposed latency.  The planner therefore peels
the last ``tail_bucket_bytes`` (default 256 KiB: the one-shot kernel's range)
off the final bucket into a bucket of their own.  For Toy_Net: {dense.*}
16.8 MB, {conv.11 ... conv.3} 6.2 MB -- launched right after conv.3's weight
gradient, overlapping conv.3's data gradient, BN0's backward and conv.0's
weight gradient -- and {conv.1, conv.0} 2.8 KB at the very end, which costs one
latency-bound one-shot kernel instead of a 6 MB transfer.

Each bucket is a contiguous slice of the flat gradient arena; slices are padded
to lcm(512 B, 32 B x world) so every bucket can be processed with 16-byte vector
/ multimem accesses (fp32 or bf16 wire) split evenly across any world size <= 8.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import torch

FIRST_BUCKET_BYTES = 1 << 20
TAIL_BUCKET_BYTES = 256 << 10
ALIGN_ELEMS = 128  # 512 B of fp32


@dataclass
class Bucket:
    index: int
    param_indices: List[int] = field(default_factory=list)   # indices into the model's param list
    offsets: List[int] = field(default_factory=list)         # element offset of each param inside the arena
    numels: List[int] = field(default_factory=list)
    start: int = 0                                            # arena element offset (aligned)
    numel: int = 0                                            # padded length in elements

    @property
    def nbytes(self) -> int:
        return sum(self.numels) * 4


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def bucket_align_elems(world_size: int = 1) -> int:
    """Bucket length granularity: 128 elements, and 8 elements (one 16-byte bf16 vector) per rank."""
    return math.lcm(ALIGN_ELEMS, 8 * max(int(world_size), 1))


def plan_buckets(params: Sequence[torch.Tensor], bucket_cap_mb: float = 25.0,
                 first_bucket_bytes: int = FIRST_BUCKET_BYTES, elem_size: int = 4,
                 world_size: int = 1, tail_bucket_bytes: int = TAIL_BUCKET_BYTES) -> List[Bucket]:
    """Assign parameters (given in *definition* order) to buckets in grad-ready order."""
    cap = int(bucket_cap_mb * 1024 * 1024)
    limits = [min(first_bucket_bytes, cap), cap]
    groups: List[List[int]] = []
    cur: List[int] = []
    cur_bytes = 0
    for idx in reversed(range(len(params))):
        cur.append(idx)
        cur_bytes += params[idx].numel() * elem_size
        limit = limits[min(len(groups), len(limits) - 1)]
        if cur_bytes >= limit:
            groups.append(cur)
            cur, cur_bytes = [], 0
    if cur:
        groups.append(cur)
    # tail split: the last-ready gradients get a small bucket of their own (exposed latency, not bandwidth)
    if tail_bucket_bytes > 0 and groups:
        last = groups[-1]
        total = sum(params[i].numel() for i in last) * elem_size
        if total > tail_bucket_bytes:
            tail_bytes, cut = 0, len(last)
            while cut > 1:
                nxt = params[last[cut - 1]].numel() * elem_size
                if tail_bytes + nxt > tail_bucket_bytes:
                    break
                tail_bytes += nxt
                cut -= 1
            if 0 < cut < len(last):
                groups[-1:] = [last[:cut], last[cut:]]
    # lay the buckets out back to back in one arena
    align = bucket_align_elems(world_size)
    buckets: List[Bucket] = []
    cursor = 0
    for g in groups:
        b = Bucket(len(buckets), param_indices=list(g), numels=[params[i].numel() for i in g])
        b.start = cursor
        inner = 0
        for n in b.numels:
            b.offsets.append(cursor + inner)
            inner += _round_up(n, 32)  # 128 B (fp32) / 64 B (bf16 shadow) alignment: TMA needs >= 16 B bases
        b.numel = _round_up(inner, align)
        cursor += b.numel
        buckets.append(b)
    return buckets


def arena_numel(buckets: Sequence[Bucket]) -> int:
    return buckets[-1].start + buckets[-1].numel if buckets else 0
