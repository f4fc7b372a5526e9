"""Host-side kernel dispatch: output bricks and which GEMM variant a layer shape is routed to (pure Python, no GPU)."""
import os

import pytest

from ddpt_b200.ops import gemm


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ("DDPT_GEMM_PAIR", "DDPT_GEMM_HALO", "DDPT_GEMM_PERSISTENT", "DDPT_WGRAD_MULTITAP"):
        monkeypatch.delenv(k, raising=False)


def test_bricks_cover_at_most_128_rows_and_waste_little():
    for nhw in ((1024, 24, 24), (1024, 10, 10), (1024, 8, 8), (304, 26, 26), (1, 1, 1024), (7, 5, 3)):
        bn, bh, bw = gemm.pick_box(*nhw)
        assert 1 <= bn * bh * bw <= 128 and bn <= nhw[0] and bh <= nhw[1] and bw <= nhw[2]
    assert gemm.pick_box(1024, 8, 8) == (2, 8, 8)                       # two whole 8x8 images per 128-row tile
    bn, bh, bw = gemm.pick_box(1024, 10, 10, 16)                        # wgrad bricks: multiples of 16 pixels
    assert (bn * bh * bw) % 16 == 0
    box, eff = gemm.pick_box_halo(1024, 24, 24)
    assert box[2] == 8 and box[0] * box[1] <= 16 and eff == 1.0        # 24 = 3 x 8: no padded columns


def test_default_routing_of_the_toy_net_layers():
    B = 1024
    mode = lambda nhw, c_k, n_total, bn: gemm._gemm_mode(None, None, nhw, 3, 3, c_k, n_total, bn)[0]
    assert mode((B, 24, 24), 64, 128, 128) == 2        # conv.3 forward: weights resident, halo reuse
    assert mode((B, 26, 26), 128, 64, 64) == 2         # conv.3 data gradient: halo reuse despite padded rows
    assert mode((B, 10, 10), 128, 256, 256) == 3       # conv.7 forward: CTA pair
    assert mode((B, 8, 8), 256, 512, 256) == 3         # conv.10 forward: CTA pair (weights do not fit)
    assert mode((B, 12, 12), 256, 128, 128) == 3       # conv.7 data gradient: CTA pair
    assert gemm._gemm_mode(None, None, (1, 1, B), 1, 1, 8192, 512, 64)[0] == 1   # dense.1 forward: 64-wide, persistent


def test_switches_override_the_defaults(monkeypatch):
    nhw = (1024, 8, 8)
    monkeypatch.setenv("DDPT_GEMM_PAIR", "0")
    assert gemm._gemm_mode(None, None, nhw, 3, 3, 256, 512, 256)[0] == 1
    monkeypatch.setenv("DDPT_GEMM_PERSISTENT", "0")
    assert gemm._gemm_mode(None, None, nhw, 3, 3, 256, 512, 256)[0] == 0
    monkeypatch.delenv("DDPT_GEMM_PERSISTENT")
    monkeypatch.setenv("DDPT_GEMM_HALO", "force")
    assert gemm._gemm_mode(None, None, nhw, 3, 3, 256, 512, 256)[0] == 2
    assert gemm._gemm_mode(True, False, nhw, 3, 3, 256, 512, 256, True)[0] == 3   # explicit arguments win
