"""C12 / N7-N8 on CPU: loss-scaler truth table and scale_loss protocol."""
import torch
import torch.nn as nn

from ddpt_b200 import amp
from ddpt_b200.amp import LossScaler
from ddpt_b200.optim import SGD, Lookahead


def test_scaler_constants_and_schedule():
    s = LossScaler(torch.device("cpu"), dynamic=True)
    assert s.loss_scale() == 2.0 ** 16
    s.begin_step(); s.unscale_([torch.tensor([float("inf")])]); s.update()
    assert s.loss_scale() == 2.0 ** 15 and s.skipped_steps() == 1
    for _ in range(1999):
        s.begin_step(); s.unscale_([torch.ones(2)]); s.update()
    assert s.loss_scale() == 2.0 ** 15
    s.begin_step(); s.unscale_([torch.ones(2)]); s.update()
    assert s.loss_scale() == 2.0 ** 16                       # doubled after 2000 clean steps
    s.load_state_dict({"scale": 2.0 ** 24, "clean_steps": 1999})
    s.begin_step(); s.unscale_([torch.ones(2)]); s.update()
    assert s.loss_scale() == 2.0 ** 24                       # capped


def test_unscale_divides_and_flags_nan():
    s = LossScaler(torch.device("cpu"), dynamic=True, init_scale=8.0)
    g = torch.tensor([8.0, 16.0])
    s.begin_step(); s.unscale_([g])
    assert g.tolist() == [1.0, 2.0] and not s.found_inf_host()
    s.begin_step(); s.unscale_([torch.tensor([float("nan")])])
    assert s.found_inf_host()


def test_scale_loss_protocol_with_forced_dynamic_scaling():
    """Overflow skips exactly one whole Lookahead step (incl. counter); clean steps match fp32."""
    torch.manual_seed(0)
    amp.reset()
    model = nn.Linear(4, 2)
    ref = nn.Linear(4, 2); ref.load_state_dict(model.state_dict())
    opt = Lookahead(SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=2, alpha=0.5)
    ropt = Lookahead(SGD(ref.parameters(), lr=0.1, momentum=0.9, nesterov=True), k=2, alpha=0.5)
    model, opt = amp.initialize(model, optimizers=opt, opt_level="O1")
    st = amp.current_state()
    assert st.opt_level == "O0"                               # CPU: plumbing config runs fp32
    st.scaler.dynamic = True; st.scaler._scale = 1024.0       # exercise dynamic path on CPU
    x, y = torch.randn(5, 4), torch.randn(5, 2)
    for step in range(4):
        opt.zero_grad(); ropt.zero_grad()
        loss = ((model(x) - y) ** 2).mean()
        with amp.scale_loss(loss, opt) as scaled:
            assert scaled.item() == loss.item() * st.scaler.loss_scale()
            scaled.backward()
            if step == 1:
                model.weight.grad[0, 0] = float("inf")       # poison -> overflow
        before = [p.detach().clone() for p in model.parameters()]
        counter = opt.param_groups[0]["counter"]
        opt.step()
        if step == 1:
            assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
            assert opt.param_groups[0]["counter"] == counter  # whole step skipped, counter too
            assert st.scaler.loss_scale() == 512.0
            continue
        ((ref(x) - y) ** 2).mean().backward(); ropt.step()
        for a, b in zip(model.parameters(), ref.parameters()):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    amp.reset()


def test_amp_disabled_is_identity():
    amp.reset()
    loss = torch.tensor(3.0, requires_grad=True)
    with amp.scale_loss(loss, []) as s:
        assert s is loss
