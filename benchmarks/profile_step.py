"""Run a few eager (non-graph) fused training steps so `ncu` sees every kernel of one step.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python benchmarks/profile_step.py --steps 1
"""
import argparse, os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpt_b200 import amp, ops
from ddpt_b200.models import Toy_Net
from ddpt_b200.optim import SGD, Lookahead

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--batch", type=int, default=1024)
a = ap.parse_args()
amp.reset()
model = Toy_Net().cuda()
opt = Lookahead(SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True), k=10, alpha=0.5)
model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
loss_fn = nn.CrossEntropyLoss()
assert ops.attach_step_engine(model, loss_fn, opt, use_graph=False)
eng = model._ddpt_step_engine
x = torch.randint(0, 256, (a.batch, 1, 28, 28), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 10, (a.batch,), device="cuda")
model.train(); eng.begin_pass()
for _ in range(3):
    eng.run(x, y, True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    eng.run(x, y, True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.steps, "steps")
