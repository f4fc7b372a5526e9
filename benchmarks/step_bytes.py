"""Whole-step DRAM-byte accounting from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log of benchmarks/profile_step.py (one eager fused Toy_Net step, batch 1024): per launch, in launch order, the
measured DRAM traffic next to the bytes the operation has to move at least once (its operands and results, bf16
activations), and the totals.

    python benchmarks/step_bytes.py gpurun_out/ncu_r2/step_bytes.csv > profiles/step_dram_bytes_r2.txt
"""
import csv, re, sys
from collections import defaultdict

B = 1024
MB = 1e6
act = {  # bf16 activation tensors of Toy_Net at batch B (bytes)
    "x": B * 28 * 28, "y0": B * 26 * 26 * 64 * 2, "y1": B * 24 * 24 * 128 * 2, "a1": B * 12 * 12 * 128 * 2,
    "y2": B * 10 * 10 * 256 * 2, "y3": B * 8 * 8 * 512 * 2, "a3": B * 4 * 4 * 512 * 2, "y4": B * 512 * 2,
}
w = {"w3": 128 * 576 * 2, "w7": 256 * 1152 * 2, "w10": 512 * 2304 * 2, "wd": 512 * 8192 * 2}
params_f32 = 23_010_856 // 4 * 4   # parameter arena bytes (5.75 M fp32)

# algorithmic bytes by (kernel, occurrence): what the op must read + write once
ALG = {
    ("conv0_fwd_kernel", 0): act["x"] + act["y0"],
    ("bn_relu_pool_kernel<0>", 0): 2 * act["y0"],                       # y0 -> a0
    ("conv_gemm_halo_kernel<128, 0, 1, 1, 1>", 0): act["y0"] + act["y1"] + w["w3"],
    ("bn_relu_pool_kernel<1>", 0): act["y1"] + act["a1"],
    ("conv_gemm_persistent_kernel<256, 0, 0, 2>", 0): act["a1"] + act["y2"] + w["w7"],
    ("bn_relu_pool_kernel<0>", 1): 2 * act["y2"],
    ("conv_gemm_persistent_kernel<256, 0, 0, 2>", 1): act["y2"] + act["y3"] + w["w10"],
    ("bn_relu_pool_flatten_kernel", 0): act["y3"] + act["a3"],
    ("chw_to_hwc_kernel", 0): 2 * act["a3"],
    ("conv_gemm_kernel<128, 0>", 0): act["a3"] + w["wd"] + 4 * B * 512 * 4,  # split-K slabs (fp32 x 4)
    ("splitk_finish_kernel", 0): 4 * B * 512 * 4 + act["y4"],
    ("bn_relu_pool_kernel<0>", 2): 2 * act["y4"],
    ("head_kernel", 0): 2 * act["y4"],
    # backward
    ("bn_bwd_reduce_kernel<0>", 0): 2 * act["y4"], ("bn_bwd_apply_kernel<0, 0>", 0): 3 * act["y4"],
    ("conv_gemm_persistent_kernel<256, 1, 0, 2>", 0): act["y4"] + w["wd"] + act["a3"],          # dense.1 dgrad
    ("wgrad_gemm_kernel<256>", 0): act["y4"] + act["a3"] + 2 * w["wd"],                           # dense.1 wgrad (fp32 out)
    ("bn_bwd_reduce_kernel<1>", 0): act["y3"] + act["a3"], ("bn_bwd_apply_kernel<1, 0>", 0): 2 * act["y3"] + act["a3"],
    ("conv_gemm_persistent_kernel<256, 1, 0, 2>", 1): act["y3"] + w["w10"] + act["y2"],         # conv.10 dgrad
    ("wgrad_multitap_kernel<128, 512>", 0): act["y3"] + act["y2"] + 2 * w["w10"],
    ("bn_bwd_reduce_kernel<0>", 1): 2 * act["y2"], ("bn_bwd_apply_kernel<0, 0>", 1): 3 * act["y2"],
    ("conv_gemm_persistent_kernel<128, 1, 0, 2>", 0): act["y2"] + w["w7"] + act["a1"],          # conv.7 dgrad
    ("wgrad_gemm_kernel<128>", 0): act["y2"] + act["a1"] + 2 * w["w7"],
    ("bn_bwd_reduce_kernel<1>", 1): act["y1"] + act["a1"], ("bn_bwd_apply_kernel<1, 0>", 1): 2 * act["y1"] + act["a1"],
    ("conv_gemm_halo_kernel<64, 1, 1, 1, 2>", 0): act["y1"] + w["w3"] + act["y0"],              # conv.3 dgrad (halo)
    ("conv_dgrad_kwmerge_kernel<2>", 0): act["y1"] + w["w3"] + act["y0"],                       # conv.3 dgrad (merged)
    ("wgrad_multitap_kernel<64, 256>", 0): act["y1"] + act["y0"] + 2 * w["w3"],
    ("bn_bwd_reduce_kernel<0>", 2): 2 * act["y0"],
    ("conv0_wgrad_kernel", 0): act["x"] + 2 * act["y0"],                 # BatchNorm-backward apply folded in: reads y0, g
    ("unscale_arena_kernel", 0): 2 * params_f32,
    ("fused_lookahead_sgd_kernel", 0): params_f32 * 3 + params_f32 * 2 + params_f32 // 2,   # p, g, momentum r/w + bf16 shadow
}


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("ddpt::<unnamed>::", "").replace("ddpt::(anonymous namespace)::", "")
    return name.replace("(int)", "").replace("(bool)", "")


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if l.startswith('"')]
    per = defaultdict(dict)
    order = []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in per:
            order.append(i); per[i]["name"] = short(r["Kernel Name"])
        per[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    seen = defaultdict(int)
    tot = defaultdict(float)
    print(f"{'#':>3} {'kernel':52s} {'us':>7} {'rd MB':>8} {'wr MB':>8} {'min MB':>8} {'x min':>6} {'GB/s':>7}")
    for i in order:
        k = per[i]; n = k["name"]; occ = seen[n]; seen[n] += 1
        us = k.get("gpu__time_duration.sum", 0) / 1e3
        rd, wr = k.get("dram__bytes_read.sum", 0) / MB, k.get("dram__bytes_write.sum", 0) / MB
        alg = ALG.get((n, occ))
        tot["us"] += us; tot["rd"] += rd; tot["wr"] += wr
        if alg is not None:
            tot["alg"] += alg / MB; tot["rdwr_model"] += rd + wr
        ratio = f"{(rd + wr) / (alg / MB):6.2f}" if alg else "     -"
        print(f"{i:3d} {n[:52]:52s} {us:7.1f} {rd:8.1f} {wr:8.1f} {(alg or 0) / MB:8.1f} {ratio} {(rd + wr) * 1e3 / max(us, 1e-9):7.0f}")
    print(f"\ntotal: {tot['us']:.1f} us serialised, DRAM read {tot['rd']:.0f} MB + write {tot['wr']:.0f} MB = {tot['rd'] + tot['wr']:.0f} MB;"
          f" modelled kernels move {tot['rdwr_model']:.0f} MB against an algorithmic minimum of {tot['alg']:.0f} MB"
          f" ({tot['rdwr_model'] / max(tot['alg'], 1e-9):.2f}x)")
    print(f"at the measured copy bandwidth the measured bytes alone take {(tot['rd'] + tot['wr']) / 6.6:.0f} us"
          f" (6.6 TB/s = 6.6 MB/us); note the 126 MB L2 lets a consumer read what the producer just wrote without DRAM traffic, so"
          f" 'x min' below 1 is real reuse, not an error")


if __name__ == "__main__":
    main()
