"""How does tcgen05.mma address 128B-swizzled operands whose start is not 1024-byte aligned?

Each experiment stages operands with TMA (absolute-address swizzle) and then points the MMA descriptors at a
start shifted by d rows (d*128 B), with the descriptor's base_offset field set to 0 or to d.  The result is
compared with the rows the shifted view *should* contain.  Needed to decide whether a convolution tile loaded
once with its halo can serve all filter taps as shifted descriptor views."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probes.build import load as _load_probes

ext = _load_probes()
torch.manual_seed(0)
dev = "cuda"
res = []


def run(name, a, b, cfg, want):
    got = ext.umma_probe(a, b, cfg)
    torch.cuda.synchronize()
    err = float((got - want).abs().max() / want.abs().max())
    res.append({"exp": name, "rel_max_err": err, "ok": err < 2e-2})
    print(f"{name:60s} rel_max_err={err:.4f} {'OK' if err < 2e-2 else 'MISMATCH'}", flush=True)


def cfg(a_rows, a_blocks, b_rows, b_blocks, a, b, ksteps):
    return [a_rows, a_blocks, b_rows, b_blocks, *a, *b, ksteps]


B64 = torch.randn(64, 64, device=dev).bfloat16()            # K-major B: [N=64, K=64]
stdB = (0, 0, 1024, 0, 0, 32)
# E1: K-major A, dense 8-row groups (SBO 1024), start shifted by d rows
A_big = torch.randn(136, 64, device=dev).bfloat16()
for d in (0, 1, 2, 5):
    want = A_big[d:d + 128].float() @ B64.float().t()
    for bo in sorted({0, d}):
        run(f"E1 Kmajor-A sbo1024 shift={d} base_offset={bo}", A_big, B64,
            cfg(136, 1, 64, 1, (d * 128, 0, 1024, bo, 0, 32), stdB, 4), want)
# E2: K-major A, 8-row groups at a 16-row pitch (SBO 2048), start shifted by d rows
A_p = torch.randn(256, 64, device=dev).bfloat16()
for d in (0, 1, 2):
    rows = torch.tensor([16 * g + d + j for g in range(16) for j in range(8)], device=dev)
    want = A_p[rows].float() @ B64.float().t()
    for bo in sorted({0, d}):
        run(f"E2 Kmajor-A sbo2048 shift={d} base_offset={bo}", A_p, B64,
            cfg(256, 1, 64, 1, (d * 128, 0, 2048, bo, 0, 32), stdB, 4), want)
# E3: MN-major B ([K rows, N=64]), K rows shifted by d
A128 = torch.randn(128, 64, device=dev).bfloat16()          # K-major A [M=128, K=64]
stdA = (0, 0, 1024, 0, 0, 32)
Bmn = torch.randn(72, 64, device=dev).bfloat16()
for d in (0, 1, 2, 5):
    want = A128.float() @ Bmn[d:d + 64].float()
    for bo in sorted({0, d}):
        run(f"E3 MNmajor-B sbo1024 kshift={d} base_offset={bo}", A128, Bmn,
            cfg(128, 1, 72, 1, stdA, (d * 128, 8192, 1024, bo, 1, 2048), 4), want)
# E4: MN-major B with K groups at a 16-row pitch
Bp = torch.randn(128, 64, device=dev).bfloat16()
for d in (0, 1, 2):
    rows = torch.tensor([16 * q + d + j for q in range(8) for j in range(8)], device=dev)
    want = A128.float() @ Bp[rows].float()
    for bo in sorted({0, d}):
        run(f"E4 MNmajor-B sbo2048 kshift={d} base_offset={bo}", A128, Bp,
            cfg(128, 1, 128, 1, stdA, (d * 128, 16384, 2048, bo, 1, 4096), 4), want)
# E5: K-major A with a 10-row pitch (SBO 1280, not a multiple of the 1024 B swizzle period)
A10 = torch.randn(176, 64, device=dev).bfloat16()
for d in (0, 1):
    rows = torch.tensor([10 * g + d + j for g in range(16) for j in range(8)], device=dev)
    want = A10[rows].float() @ B64.float().t()
    for bo in sorted({0, d}):
        run(f"E5 Kmajor-A sbo1280 shift={d} base_offset={bo}", A10, B64,
            cfg(176, 1, 64, 1, (d * 128, 0, 1280, bo, 0, 32), stdB, 4), want)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_umma_desc.json", "w"), indent=1)
