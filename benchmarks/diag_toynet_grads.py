"""Per-parameter gradient error of the fused Toy_Net backward vs the fp32 PyTorch graph."""
import copy, os, sys
import torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpt_b200.models import Toy_Net
from ddpt_b200.ops import toynet_fused
from ddpt_b200.utils import same_seeds

same_seeds(3)
ref = Toy_Net().cuda()
with torch.no_grad():
    for m in ref.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
            m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
fused = copy.deepcopy(ref)
toynet_fused.attach(fused, "bf16")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randint(0, 256, (128, 1, 28, 28), device="cuda", generator=g, dtype=torch.uint8)
y = torch.randint(0, 10, (128,), device="cuda", generator=g)
# a bf16-autocast torch run shows how much error low precision alone introduces
auto = copy.deepcopy(ref)
ref.train(); fused.train(); auto.train()
F.cross_entropy(ref.forward_reference(x), y).backward()
F.cross_entropy(fused(x), y).backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = auto.forward_reference(x)
F.cross_entropy(out.float(), y).backward()
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))
print(f"{'param':18s} {'fused_vs_fp32':>14s} {'autocast_vs_fp32':>17s} {'|g|_fp32':>10s}")
for (n, pr), pf, pa in zip(ref.named_parameters(), fused.parameters(), auto.parameters()):
    print(f"{n:18s} {rel(pf.grad, pr.grad):14.4f} {rel(pa.grad, pr.grad):17.4f} {float(pr.grad.norm()):10.3e}")
