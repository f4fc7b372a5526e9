"""Gradient bucket planning (what DDP's ``_compute_bucket_assignment_by_size`` +
the one-time rebuild converge to).

Gradients become ready in roughly reverse definition order, so buckets are
planned over the *reversed* parameter list with a small first bucket (1 MiB) so
communication starts early, then ``bucket_cap_mb`` buckets.  For Toy_Net this
yields the reference's steady-state layout directly -- {dense.4, dense.2,
dense.1} = 16 803 880 B, then {all conv/BN2d} = 6 206 976 B -- with no rebuild
iteration and no bucket-index broadcast (N6): the plan is a pure function of the
parameter shapes, hence identical on every rank.

Each bucket is a contiguous slice of the flat gradient arena; slices are padded
to 512 B so every bucket can be processed with 16-byte vector / multimem
accesses split evenly across up to 8 ranks.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import torch

FIRST_BUCKET_BYTES = 1 << 20
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


def plan_buckets(params: Sequence[torch.Tensor], bucket_cap_mb: float = 25.0,
                 first_bucket_bytes: int = FIRST_BUCKET_BYTES, elem_size: int = 4) -> List[Bucket]:
    """Assign parameters (given in *definition* order) to buckets in grad-ready order."""
    cap = int(bucket_cap_mb * 1024 * 1024)
    limits = [min(first_bucket_bytes, cap), cap]
    buckets: List[Bucket] = []
    cur = Bucket(0)
    cur_bytes = 0
    for idx in reversed(range(len(params))):
        n = params[idx].numel()
        cur.param_indices.append(idx)
        cur.numels.append(n)
        cur_bytes += n * elem_size
        limit = limits[min(len(buckets), len(limits) - 1)]
        if cur_bytes >= limit:
            buckets.append(cur)
            cur, cur_bytes = Bucket(len(buckets)), 0
    if cur.param_indices:
        buckets.append(cur)
    # lay the buckets out back to back in one arena
    cursor = 0
    for b in buckets:
        b.start = cursor
        inner = 0
        for n in b.numels:
            b.offsets.append(cursor + inner)
            inner += _round_up(n, 32)  # 128 B (fp32) / 64 B (bf16 shadow) alignment: TMA needs >= 16 B bases
        b.numel = _round_up(inner, ALIGN_ELEMS)
        cursor += b.numel
    return buckets


def arena_numel(buckets: Sequence[Bucket]) -> int:
    return buckets[-1].start + buckets[-1].numel if buckets else 0
