"""Time Toy_Net conv.3's data gradient (batch 1024): merged-column kernel vs the halo kernel vs cuDNN."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpt_b200.ops import gemm

def timeit(fn, iters=50):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dy = torch.randn(n, 24, 24, 128, device="cuda").bfloat16()
wk = (torch.randn(128, 576, device="cuda") * 0.03).bfloat16()
out = torch.empty(n, 26, 26, 64, device="cuda", dtype=torch.bfloat16)
res = {"kwmerge_us": timeit(lambda: gemm.conv_dgrad(dy, wk, 3, 3, out=out, kwmerge=True)),
       "halo_us": timeit(lambda: gemm.conv_dgrad(dy, wk, 3, 3, out=out, kwmerge=False))}
wt = wk.view(128, 3, 3, 64).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
dyc = dy.permute(0, 3, 1, 2)
res["cudnn_us"] = timeit(lambda: torch.ops.aten.convolution_backward(dyc, torch.empty(n, 64, 26, 26, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last), wt, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False]))
flops = 2.0 * n * 26 * 26 * 64 * 1152
res = {k: round(v, 1) for k, v in res.items()}
res["kwmerge_tflops"] = round(flops / res["kwmerge_us"] / 1e6, 1)
print(res)
