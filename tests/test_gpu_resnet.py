"""Generic-model path on the GPU: tensor-core pointwise convolutions behind autograd and the CUDA-graphed step."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,c_in,c_out,hw,stride", [(4, 64, 256, 14, 1), (3, 256, 64, 9, 1), (2, 128, 512, 8, 2),
                                                    (64, 64, 64, 56, 1)])
def test_tensor_core_conv1x1_matches_conv2d(n, c_in, c_out, hw, stride):
    from ddpt_b200.ops.conv_modules import TensorCoreConv1x1
    torch.manual_seed(0)
    ours = TensorCoreConv1x1(c_in, c_out, 1, stride=stride, bias=False).cuda()
    ref = nn.Conv2d(c_in, c_out, 1, stride=stride, bias=False).cuda()
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(n, c_in, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last)
    xo, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    before = TensorCoreConv1x1.fused_calls
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yo = ours(xo)
        yr = ref(xr)
    assert TensorCoreConv1x1.fused_calls == before + 1, "the tcgen05 path did not run"
    assert yo.shape == yr.shape and yo.is_contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(yr.float())
    yo.float().backward(g); yr.float().backward(g)
    err = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))
    assert err(yo, yr) < 1e-2
    assert err(xo.grad, xr.grad) < 1e-2
    assert err(ours.weight.grad, ref.weight.grad) < 1e-2


class _SmallNet(nn.Module):
    def __init__(self):
        super().__init__()
        from ddpt_b200.ops.conv_modules import TensorCoreConv1x1
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 3, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU())
        self.pw1 = TensorCoreConv1x1(64, 128, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(128)
        self.pw2 = TensorCoreConv1x1(128, 64, 1, stride=2, bias=False)
        self.head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(64, 10))

    def forward(self, x):
        x = self.stem(x)
        x = F.relu(self.bn1(self.pw1(x)))
        return self.head(F.relu(self.pw2(x)))


def test_graphed_train_step_matches_eager_and_trains():
    """GraphedTrainStep (forward + backward + loss scaling + fused Lookahead-SGD as one CUDA graph) follows the same
    trajectory as the eager loop of utils/process.py:17-28 and the loss goes down."""
    from ddpt_b200 import amp
    from ddpt_b200.engine import GraphedTrainStep
    from ddpt_b200.ops.conv_modules import use_tensor_core_pointwise_convs
    from ddpt_b200.optim import SGD, Lookahead
    from ddpt_b200.utils import same_seeds
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(32, 3, 16, 16, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
          for _ in range(12)]
    ys = [(x.mean((1, 2, 3)) > 0).long() + 2 * (x[:, 0, 0, 0] > 0).long() for x in xs]
    loss_fn = nn.CrossEntropyLoss()
    losses = {}
    for mode in ("graph", "eager"):
        amp.reset()
        same_seeds(3)
        model = _SmallNet().cuda().to(memory_format=torch.channels_last)
        assert use_tensor_core_pointwise_convs(model) == 2
        opt = Lookahead(SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True), k=4, alpha=0.5)
        model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
        step = GraphedTrainStep(model, loss_fn, opt, xs[0], ys[0], warmup=2, use_graph=(mode == "graph"))
        out = [float(step(x, y)) for x, y in zip(xs, ys)]
        assert (step.replays == len(xs)) == (mode == "graph")
        losses[mode] = out
        amp.reset()
    assert losses["graph"][-1] < 0.7 * losses["graph"][0]
    for a, b in zip(losses["graph"], losses["eager"]):
        assert a == pytest.approx(b, rel=5e-2, abs=5e-3)
