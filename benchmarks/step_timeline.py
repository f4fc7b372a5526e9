"""Kernel timeline of ONE CUDA-graph replay of the fused Toy_Net step (CUPTI through torch.profiler): start offset,
duration and stream of every kernel, so overlap between the main stream and the weight-gradient side stream -- and the
critical path -- can be read off.  (Profiler timestamps; durations under CUPTI run a little long.)

    python benchmarks/step_timeline.py [--batch 1024] > profiles/step_timeline_r2.txt
"""
import argparse, os, re, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpt_b200 import amp, ops
from ddpt_b200.models import Toy_Net
from ddpt_b200.optim import SGD, Lookahead
from torch.profiler import ProfilerActivity, profile

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
a = ap.parse_args()
amp.reset()
model = Toy_Net().cuda()
opt = Lookahead(SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True), k=10, alpha=0.5)
model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
assert ops.attach_step_engine(model, nn.CrossEntropyLoss(), opt, use_graph=True)
eng = model._ddpt_step_engine
x = torch.randint(0, 256, (a.batch, 1, 28, 28), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 10, (a.batch,), device="cuda")
model.train(); eng.begin_pass()
for _ in range(8):
    eng.run(x, y, True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        eng.run(x, y, True)
    torch.cuda.synchronize()
def short(n):
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("ddpt::", "")
    return n.replace("at::native::", "")[:60]


acts = []          # (start us, duration us, stream, name)
try:
    for e in prof.profiler.kineto_results.events():
        if "cuda" in str(e.device_type()).lower() or "CUDA" in str(e.device_type()):
            acts.append((e.start_ns() / 1e3, e.duration_ns() / 1e3, e.device_resource_id(), e.name()))
except Exception as err:                     # older / newer kineto bindings: fall back to the public events (no stream id)
    print("kineto_results unavailable:", err)
    for e in prof.events():
        if e.device_type is not None and "cuda" in str(e.device_type).lower():
            acts.append((e.time_range.start, e.time_range.end - e.time_range.start, 0, e.name))
acts.sort()
per = len(acts) // 3
last = acts[-per:]                           # the last of the three replays
t0 = last[0][0]
sid = {s: i for i, s in enumerate(sorted({a[2] for a in last}))}
print(f"{len(last)} device activities in one replay; columns: start us, end us, duration us, stream, kernel")
print("(a kernel launched with programmatic dependent launch 'starts' when its CTAs are scheduled and then waits for its")
print(" predecessor: overlapping rows on ONE stream are that wait, rows on different streams are real concurrency)")
busy_end, gap_total = 0.0, 0.0
for s0, d, st, name in last:
    s = s0 - t0
    gap = s - busy_end
    if gap > 0 and busy_end > 0:
        gap_total += gap
    busy_end = max(busy_end, s + d)
    print(f"{s:9.1f} {s + d:9.1f} {d:7.1f}  s{sid[st]}  {short(name)}" + (f"   (device idle {gap:.1f})" if gap > 1.0 and s > 0 else ""))
print(f"replay span {busy_end:.1f} us, sum of kernel durations {sum(a[1] for a in last):.1f} us, "
      f"device idle inside the span {gap_total:.1f} us")
