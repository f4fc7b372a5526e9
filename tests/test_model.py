"""C5/C17/C16: Toy_Net layout + init parity, checkpoint layout, TensorBoard layout, buckets."""
import os

import pytest
import torch

from conftest import load_reference_module
from ddpt_b200.models import Toy_Net, resnet50, toy_net_flops_per_image
from ddpt_b200.parallel import plan_buckets
from ddpt_b200.utils import ScalarLogger, load_weights, same_seeds, save_weights


def test_param_and_buffer_inventory():
    m = Toy_Net()
    assert sum(p.numel() for p in m.parameters()) == 5_752_714
    assert len(list(m.parameters())) == 22 and len(list(m.buffers())) == 15
    assert len(m.state_dict()) == 37
    assert m(torch.rand(3, 1, 28, 28)).shape == (3, 10)
    assert toy_net_flops_per_image() == pytest.approx(304.1e6, rel=2e-3)


def test_uint8_input_equals_totensor():
    m = Toy_Net().eval()
    u8 = torch.randint(0, 256, (4, 1, 28, 28), dtype=torch.uint8)
    torch.testing.assert_close(m(u8), m(u8.float() / 255.0))


def test_init_bit_identical_to_reference(reference_root):
    ref = load_reference_module("model")
    same_seeds(42); ours = Toy_Net()
    same_seeds(42); theirs = ref.Toy_Net()
    a, b = ours.state_dict(), theirs.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_checkpoint_layout_matches_shipped_artifact(reference_root, tmp_path):
    shipped = torch.load(os.path.join(reference_root, "experiment_model/DDP_warmup.pt"),
                         map_location="cpu", weights_only=True)
    m = Toy_Net()
    path = str(tmp_path / "exp.pt")
    save_weights(m, path)
    ours = torch.load(path, map_location="cpu", weights_only=True)
    assert list(ours) == list(shipped)
    for k in shipped:
        assert ours[k].shape == shipped[k].shape and ours[k].dtype == shipped[k].dtype, k
    # both directions load strictly
    ref = load_reference_module("model").Toy_Net()
    ref.load_state_dict(ours, strict=True)
    load_weights(m, os.path.join(reference_root, "experiment_model/DDP_warmup.pt"))
    assert int(m.state_dict()["conv.1.num_batches_tracked"]) == 14790


def test_bucket_plan_matches_ddp_steady_state():
    m = Toy_Net()
    params = list(m.parameters())
    # torch DDP's steady-state layout when the tail split is off ...
    buckets = plan_buckets(params, 25.0, tail_bucket_bytes=0)
    assert [b.nbytes for b in buckets] == [16_803_880, 6_206_976]
    names = [n for n, _ in m.named_parameters()]
    assert [names[i] for i in buckets[0].param_indices] == [
        "dense.4.bias", "dense.4.weight", "dense.2.bias", "dense.2.weight", "dense.1.bias", "dense.1.weight"]
    assert all(o % 4 == 0 for b in buckets for o in b.offsets) and all(b.numel % 128 == 0 for b in buckets)
    # ... and by default the last-ready gradients (conv.1 / conv.0: nothing left to overlap with) get their own
    # latency-sized bucket; everything else keeps DDP's order
    buckets = plan_buckets(params, 25.0)
    assert [b.nbytes for b in buckets] == [16_803_880, 6_203_904, 3_072]
    assert [names[i] for i in buckets[2].param_indices] == ["conv.1.bias", "conv.1.weight", "conv.0.bias", "conv.0.weight"]
    assert sorted(i for b in buckets for i in b.param_indices) == list(range(22))
    for world in (2, 3, 5, 6, 7, 8):        # every bucket splits into 16-byte vectors (fp32 or bf16) per rank
        plan = plan_buckets(params, 25.0, world_size=world)
        assert all(b.numel % (8 * world) == 0 and b.numel % 128 == 0 and b.start % 128 == 0 for b in plan)
    big = plan_buckets(list(resnet50().parameters()), 25.0)
    assert len(big) >= 4 and all(b.nbytes <= 36 * 2**20 for b in big[1:]) and big[-1].nbytes <= 256 * 1024


def test_tensorboard_layout(tmp_path):
    log = ScalarLogger(str(tmp_path / "logs" / "exp"))
    for e in range(3):
        log.epoch(e, 0.1 * e, 1.0, 2.0, 0.5, 0.6)
    log.close()
    root = tmp_path / "logs" / "exp"
    subs = sorted(p.name for p in root.iterdir() if p.is_dir())
    assert subs == ["Acc_train", "Acc_valid", "Loss_train", "Loss_valid"]
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(str(root)); acc.Reload()
    assert acc.Tags()["scalars"] == ["LR"]
    assert [s.value for s in acc.Scalars("LR")] == pytest.approx([0.0, 0.1, 0.2])
    sub = EventAccumulator(str(root / "Loss_valid")); sub.Reload()
    assert sub.Tags()["scalars"] == ["Loss"] and len(sub.Scalars("Loss")) == 3
