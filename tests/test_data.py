"""C3: sampler parity with torch DistributedSampler, IDX reader, batch loader tails."""
import os

import pytest
import torch
from torch.utils.data.distributed import DistributedSampler

from ddpt_b200.data import BatchLoader, ShardedSampler, load_mnist, synthetic_mnist


@pytest.mark.parametrize("n,world", [(60000, 2), (60000, 8), (1001, 4), (10, 3)])
@pytest.mark.parametrize("epoch", [0, 3])
def test_sampler_matches_torch(n, world, epoch):
    ds = range(n)
    for rank in range(world):
        ref = DistributedSampler(ds, num_replicas=world, rank=rank)
        ref.set_epoch(epoch)
        ours = ShardedSampler(n, num_replicas=world, rank=rank)
        ours.set_epoch(epoch)
        assert list(ref) == ours.indices().tolist()
        assert len(ref) == len(ours)


def test_sampler_no_shuffle_and_drop_last():
    for kw in (dict(shuffle=False), dict(drop_last=True), dict(shuffle=False, drop_last=True)):
        ref = DistributedSampler(range(103), num_replicas=4, rank=1, **kw)
        ours = ShardedSampler(103, num_replicas=4, rank=1, **kw)
        assert list(ref) == ours.indices().tolist()


def test_tail_batches_like_reference():
    data = synthetic_mnist(60000)
    for world, steps, tail in [(1, 59, 608), (2, 30, 304), (4, 15, 664), (8, 8, 332)]:
        loader = BatchLoader(data, 1024, ShardedSampler(60000, world, 0), mode="torch")
        assert len(loader) == steps
        sizes = [img.shape[0] for img, _ in loader]
        assert sizes[-1] == tail and all(s == 1024 for s in sizes[:-1])


def test_loader_values_are_totensor_equivalent():
    data = synthetic_mnist(300)
    loader = BatchLoader(data, 128, None, mode="torch")
    img, tgt = next(iter(loader))
    assert img.shape == (128, 1, 28, 28) and img.dtype == torch.float32 and tgt.dtype == torch.int64
    torch.testing.assert_close(img[:, 0], data.images[:128].float() / 255.0, rtol=0, atol=0)
    assert torch.equal(tgt, data.labels[:128])


def test_idx_reader_on_reference_dataset(reference_root):
    root = os.path.join(reference_root, "datasets")
    train, test = load_mnist(root, True), load_mnist(root, False)
    assert train.images.shape == (60000, 28, 28) and train.images.dtype == torch.uint8
    assert test.images.shape == (10000, 28, 28) and train.labels[:5].tolist() == [5, 0, 4, 1, 9]
    processed = torch.load(os.path.join(root, "MNIST/processed/training.pt"), weights_only=True)
    assert torch.equal(processed[0], train.images) and torch.equal(processed[1], train.labels)


def test_missing_dataset_message(tmp_path):
    with pytest.raises(FileNotFoundError, match="--synthetic"):
        load_mnist(str(tmp_path))


def test_native_threaded_loader_matches_indexing():
    """csrc/data/host_loader.cpp: threaded gather into rotating buffers (pinned on CUDA boxes)."""
    from ddpt_b200 import ops
    if ops.load_ext() is None:
        pytest.skip("extension not built")
    from ddpt_b200.ops.host_loader import PinnedGatherLoader
    data = synthetic_mnist(3000, seed=4)
    loader = PinnedGatherLoader(data.images, data.labels, 256, None, depth=4, threads=3)
    sampler = ShardedSampler(3000, 2, 1)
    for epoch in range(3):
        sampler.set_epoch(epoch)
        idx = sampler.indices()
        nb = -(-len(idx) // 256)
        imgs, lbls, sizes = [], [], []
        for img, lbl, nbytes in loader.epoch(idx, nb):
            imgs.append(img.clone()); lbls.append(lbl.clone()); sizes.append(img.shape[0])
        assert sizes == [256] * (nb - 1) + [len(idx) - 256 * (nb - 1)]
        assert torch.equal(torch.cat(imgs)[:, 0], data.images[idx]) and torch.equal(torch.cat(lbls), data.labels[idx])


def test_native_loader_speculative_next_epoch_and_restart():
    """The native gather loader starts the NEXT epoch's batches as soon as the current epoch's last batch is out; the
    next call either adopts that work (same indices) or discards it (different indices) -- same batches either way."""
    from ddpt_b200.ops import load_ext
    if load_ext() is None:
        pytest.skip("extension not built")
    from ddpt_b200.ops.host_loader import PinnedGatherLoader
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (500, 28, 28), dtype=torch.uint8, generator=g)
    labels = torch.randint(0, 10, (500,), generator=g)
    loader = PinnedGatherLoader(images, labels, 64, device=None, depth=4, threads=2)
    perms = [torch.randperm(500, generator=g) for _ in range(4)]

    def run(idx, nxt=None):
        out = []
        for img, lbl, _ in loader.epoch(idx, 8, (lambda: (nxt, 8)) if nxt is not None else None):
            out.append((img.clone(), lbl.clone()))
        return out

    def check(batches, idx):
        assert len(batches) == 8
        for b, (img, lbl) in enumerate(batches):
            sel = idx[b * 64:(b + 1) * 64]
            assert torch.equal(img.view(-1, 28, 28), images[sel]) and torch.equal(lbl, labels[sel])

    check(run(perms[0], perms[1]), perms[0])
    assert (loader.epoch_starts, loader.adopted_epochs) == (2, 0)        # epoch 0 + the speculative start of epoch 1
    check(run(perms[1], perms[2]), perms[1])                  # adopted: no restart, the head of the epoch was already gathered
    assert (loader.epoch_starts, loader.adopted_epochs) == (3, 1)        # only epoch 2's speculation was started
    check(run(perms[3]), perms[3])                            # speculation (perms[2]) discarded, fresh start
    assert (loader.epoch_starts, loader.adopted_epochs) == (4, 1)
    check(run(perms[0]), perms[0])
    assert (loader.epoch_starts, loader.adopted_epochs) == (5, 1)
    # every batch of the four epochs was gathered; speculative batches of the discarded epoch come on top (at most `depth`)
    assert 32 <= loader._native.batches_gathered() <= 32 + 4
