"""Launch one Toy_Net convolution GEMM a few times (for `ncu -k regex:... -s 2 -c 1`).

    python benchmarks/run_one_conv.py conv.10 fwd pair | default
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpt_b200.ops import gemm  # noqa: E402

SHAPES = {"conv.3": (26, 26, 64, 128), "conv.7": (12, 12, 128, 256), "conv.10": (10, 10, 256, 512)}


def main():
    layer, op, flavour = sys.argv[1], sys.argv[2], sys.argv[3]
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    h, w, ci, co = SHAPES[layer]
    x = torch.randn(batch, h, w, ci, device="cuda").bfloat16()
    wk = (torch.randn(co, 9 * ci, device="cuda") * 0.05).bfloat16()
    dy = torch.randn(batch, h - 2, w - 2, co, device="cuda").bfloat16()
    kw = dict(persistent=flavour != "per_tile", halo=flavour == "halo", pair=flavour == "pair")
    if flavour == "default":        # what the fused model's step launches
        kw = {}
    for _ in range(4):
        if op == "fwd":
            gemm.conv_fwd(x, wk, 3, 3, **kw)
        elif op == "dgrad":
            gemm.conv_dgrad(dy, wk, 3, 3, **kw)
        else:
            gemm.conv_wgrad(dy, x, 3, 3, multitap=flavour == "multitap")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
