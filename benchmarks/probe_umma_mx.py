"""Where do the UE8M0 scale factors of tcgen05.mma.kind::mxf8f6f4.block_scale live in TMEM?

Operands are all-ones e4m3 (0x38), so D[m][n] = sum over the four K=32 MMAs of 32 * 2^(sfa-127) * 2^(sfb-127);
scale words are written to TMEM lane-by-lane with tcgen05.st, and the pattern of D reveals which
(lane, column, byte) feeds which (row, k-block)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probes.build import load as _load_probes
ext = _load_probes()
dev = "cuda"
ONE = 0x38
A = torch.full((128, 128), ONE, dtype=torch.uint8, device=dev)
B = torch.full((64, 128), ONE, dtype=torch.uint8, device=dev)


def words(fn):
    """[128 lanes][4 cols] int32 from fn(lane, col, byte) -> exponent byte."""
    w = torch.zeros(128, 4, dtype=torch.int64)
    for lane in range(128):
        for col in range(4):
            v = 0
            for byte in range(4):
                v |= (fn(lane, col, byte) & 0xFF) << (8 * byte)
            w[lane, col] = v if v < 2 ** 31 else v - 2 ** 32
    return w.to(torch.int32).to(dev)


unit = words(lambda l, c, b: 127)
log = []


def run(name, sfa, sfb, mode=1, ksteps=4, sfa_col=64, sfb_col=96):
    try:
        d = ext.umma_mx_probe(A, B, sfa, sfb, sfa_col, sfb_col, mode, ksteps)
        torch.cuda.synchronize()
    except Exception as e:
        print(name, "FAILED", repr(e)[:200]); log.append({"exp": name, "error": repr(e)[:200]}); return None
    rows = d[:, 0].tolist()
    cols = d[0, :].tolist()
    print(f"{name}\n   D[0:10,0]={[round(x, 2) for x in rows[:10]]} D[32:36,0]={[round(x,2) for x in rows[32:36]]} "
          f"D[64:66,0]={[round(x,2) for x in rows[64:66]]} D[96:98,0]={[round(x,2) for x in rows[96:98]]}\n"
          f"   D[0,0:10]={[round(x, 2) for x in cols[:10]]} D[0,32:36]={[round(x,2) for x in cols[32:36]]}", flush=True)
    log.append({"exp": name, "col0": rows, "row0": cols})
    return d


run("E1 all scales 1.0 (expect 128)", unit, unit)
run("E1b one MMA only (expect 32)", unit, unit, ksteps=1)
byte_ramp = words(lambda l, c, b: 127 + b if c == 0 else 127)
run("E2 sfa col0 byte j = 2^j, sf_id=k (480 if MMA k uses byte k; 128 if byte 0)", byte_ramp, unit, mode=1)
run("E2b same, sf_id=0 for all MMAs", byte_ramp, unit, mode=0)
lane_ramp = words(lambda l, c, b: 127 + (l % 8) if c == 0 else 127)
run("E3 sfa col0 = 2^(lane%8): row m <-> lane m ?", lane_ramp, unit)
col_probe = words(lambda l, c, b: 127 if c == 0 else 127 + 3 + c)
run("E4 sfa cols 1..3 = 2^(3+c): are other columns read?", col_probe, unit)
run("E5 sfb col0 = 2^(lane%8): column n <-> lane n ?", unit, lane_ramp)
run("E5b sfb byte ramp, sf_id=k", unit, byte_ramp, mode=1)
quad_probe = words(lambda l, c, b: 127 + (l // 32) if c == 0 else 127)
run("E6 sfa col0 = 2^(lane//32): quadrant mapping", quad_probe, unit)
run("E6b sfb col0 = 2^(lane//32)", unit, quad_probe)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(log, open("gpurun_out/probe_umma_mx.json", "w"))
