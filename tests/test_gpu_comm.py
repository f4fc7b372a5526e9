"""Fused all-reduce kernel: protocol + numerics.

Single-GPU tests simulate N ranks with N buffers and N concurrently resident kernels on N streams
(the cross-GPU barrier only needs the CTAs to be co-resident); the multimem (NVLS) variants need a real
multicast mapping and are covered by tests/test_gpu_multi.py (multi-GPU only, skipped with fewer than
2 devices)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

ONESHOT, TWOSHOT = 1, 2
F32, BF16 = 0, 1


@pytest.fixture(scope="module")
def ext():
    from ddpt_b200 import ops
    return ops.require_native()


class Sim:
    """`world` simulated ranks on one GPU: arenas, signal pads, scratch regions, one stream each."""

    def __init__(self, ext, world, total):
        dev = torch.device("cuda")
        self.ext, self.world, self.total = ext, world, total
        self.bufs = [torch.zeros(total, device=dev) for _ in range(world)]
        self.sigs = [torch.zeros(ext.signal_pad_bytes() // 4, dtype=torch.int32, device=dev) for _ in range(world)]
        self.scratch = [torch.zeros(ext.scratch_bytes(total), dtype=torch.uint8, device=dev) for _ in range(world)]
        self.streams = [torch.cuda.Stream() for _ in range(world)]

    def fill(self, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        for b in self.bufs:
            b.copy_(torch.randn(self.total, device="cuda", generator=g))

    def run(self, offset, numel, algo, wire, states=None, channel=3):
        torch.cuda.synchronize()
        for r in range(self.world):
            with torch.cuda.stream(self.streams[r]):
                st = states[r] if states is not None else None
                self.ext.fused_allreduce([b.data_ptr() for b in self.bufs], [s.data_ptr() for s in self.sigs], 0,
                                         [s.data_ptr() for s in self.scratch], 0, r, offset, numel, st,
                                         st[2:3] if st is not None else None, channel, algo, wire, 0, 4000)
        torch.cuda.synchronize()


def _check(ext, world, numel, algo, wire, scale_state=False, poison=None, iters=1, offset=256):
    sim = Sim(ext, world, offset + numel + 64)
    states = ([torch.tensor([4.0, 0.25, 0.0, 0.0, 0.0], device="cuda") for _ in range(world)] if scale_state else None)
    for it in range(iters):
        sim.fill(world * 1000 + numel % 997 + it)
        if poison is not None:
            sim.bufs[poison[0]][offset + poison[1]] = float("inf")
        srcs = [b[offset:offset + numel].clone() for b in sim.bufs]
        if wire == BF16:                                             # what the wire carries
            srcs = [s.bfloat16().float() for s in srcs]
        want = torch.stack(srcs).double().sum(0) * ((0.25 if scale_state else 1.0) / world)
        untouched = [b.clone() for b in sim.bufs]
        sim.run(offset, numel, algo, wire, states)
        for r in range(world):
            got = sim.bufs[r][offset:offset + numel]
            if poison is None:
                if wire == BF16 and (algo == TWOSHOT or (algo == 0 and numel > 65536)):   # gathered result is bf16 too
                    torch.testing.assert_close(got.double(), want, rtol=2 ** -8, atol=1e-6)
                else:
                    torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-5)
            assert torch.equal(sim.bufs[r][:offset], untouched[r][:offset])
            assert torch.equal(sim.bufs[r][offset + numel:], untouched[r][offset + numel:])
        for r in range(1, world):                                    # every rank holds the same bits
            a, b = sim.bufs[0][offset:offset + numel], sim.bufs[r][offset:offset + numel]
            assert torch.equal(a.view(torch.int32), b.view(torch.int32)), f"rank {r} differs from rank 0 (iter {it})"
    return sim, states


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("wire", [F32, BF16])
@pytest.mark.parametrize("algo,numel", [(ONESHOT, 1920), (ONESHOT, 65280), (TWOSHOT, 3840),
                                        (TWOSHOT, 16 * 1024 * 1024 // 4 + 1856)])
def test_simulated_ranks_reduce_in_place(ext, world, wire, algo, numel):
    if world == 8 and numel > 1 << 20:
        pytest.skip("8 x 32 co-resident CTAs of the large bucket are covered on real GPUs")
    _check(ext, world, numel, algo, wire)


@pytest.mark.parametrize("wire", [F32, BF16])
def test_all_ranks_bit_identical_and_scaled(ext, wire):
    sim, states = _check(ext, 4, 8192 * 8, TWOSHOT, wire, scale_state=True, offset=128)
    assert all(float(s[2]) == 0.0 for s in states)
    sim, states = _check(ext, 3, 3 * 8 * 100, ONESHOT, wire, scale_state=True, offset=128)
    assert all(float(s[2]) == 0.0 for s in states)


@pytest.mark.parametrize("wire", [F32, BF16])
@pytest.mark.parametrize("algo", [ONESHOT, TWOSHOT])
def test_overflow_reaches_every_rank(ext, algo, wire):
    # the poisoned element sits in the LAST block's share: with several CTAs per rank this is the case the old
    # single-word protocol could miss (block 0 collected before block k published)
    numel = 4096 if algo == ONESHOT else 8 * 1024 * 1024 // 4
    sim, states = _check(ext, 4, numel, algo, wire, scale_state=True, poison=(2, numel - 5), offset=0)
    assert all(float(s[2]) == 1.0 for s in states), [float(s[2]) for s in states]
    # and the flag does not leak into the next (clean) launch
    for s in states:
        s[2] = 0.0
    sim.fill(5)
    sim.run(0, numel, algo, wire, states)
    assert all(float(s[2]) == 0.0 for s in states), [float(s[2]) for s in states]


@pytest.mark.parametrize("wire", [F32, BF16])
def test_back_to_back_iterations_reuse_flags(ext, wire):
    _check(ext, 4, 65536, TWOSHOT, wire, iters=6)
    _check(ext, 4, 2048, ONESHOT, wire, iters=7)          # odd count: both push-slot parities, twice


def test_mixed_sizes_share_a_channel(ext):
    """Buckets of different sizes (different grids) alternate on one channel: per-block epochs stay consistent."""
    sim = Sim(ext, 4, 1 << 20)
    for it, (numel, algo) in enumerate([(512, ONESHOT), (65536, ONESHOT), (1 << 19, TWOSHOT), (2048, ONESHOT),
                                        (65536, ONESHOT), (4096, TWOSHOT), (512, ONESHOT)] * 2):
        sim.fill(it)
        want = torch.stack([b[:numel] for b in sim.bufs]).double().sum(0) / 4
        sim.run(0, numel, algo, F32)
        for b in sim.bufs:
            torch.testing.assert_close(b[:numel].double(), want, rtol=1e-5, atol=1e-5)


def test_auto_picks_by_size(ext):
    _check(ext, 2, 1024, 0, F32)
    _check(ext, 2, 1 << 20, 0, F32)
    _check(ext, 2, 1 << 20, 0, BF16)


def test_symm_broadcast_simulated(ext):
    world, n = 4, 3000
    dev = torch.device("cuda")
    bufs = [torch.full((n + 8,), float(r), device=dev) for r in range(world)]
    sigs = [torch.zeros(ext.signal_pad_bytes() // 4, dtype=torch.int32, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it in range(3):
        bufs[1][4:4 + n] = torch.arange(n, device=dev, dtype=torch.float32) + it
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ext.symm_broadcast([b.data_ptr() for b in bufs], [s.data_ptr() for s in sigs], 0, r, 1, 4, n, 7, 4000)
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r][4:4 + n], bufs[1][4:4 + n])
            assert float(bufs[r][0]) == float(r) and float(bufs[r][-1]) == float(r)


def test_bad_arguments_raise(ext):
    sim = Sim(ext, 2, 4096)
    with pytest.raises(RuntimeError):                      # two-shot bucket not divisible by 16 B x world
        sim.ext.fused_allreduce([b.data_ptr() for b in sim.bufs], [s.data_ptr() for s in sim.sigs], 0,
                                [s.data_ptr() for s in sim.scratch], 0, 0, 0, 1028, None, None, 0, TWOSHOT, F32, 0, 100)
    with pytest.raises(RuntimeError):                      # bf16 wire without a scratch region
        sim.ext.fused_allreduce([b.data_ptr() for b in sim.bufs], [s.data_ptr() for s in sim.sigs], 0,
                                [], 0, 0, 0, 1024, None, None, 0, TWOSHOT, BF16, 0, 100)
