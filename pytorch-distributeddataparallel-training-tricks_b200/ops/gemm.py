"""Tensor-core convolution / linear primitives (tcgen05 + TMEM + TMA), NHWC bf16 (or fp16: same kernels, the
instruction descriptor's operand format and the epilogue's pack follow the tensors' dtype).

Thin wrappers over ``_ddpt_C.conv_gemm`` / ``_ddpt_C.wgrad_gemm`` that pick the output brick
(the M tile is a box of output pixels, see csrc/nn/conv_gemm.cu) and the N tile.  Layout
conventions used throughout the fused Toy_Net path:

* activations: ``[N, H, W, C]`` bf16 contiguous;
* conv weights: ``[C_out, kh*kw*C_in]`` bf16 -- the channels-last physical order of the fp32
  master weights in the parameter arena, so the bf16 shadow is a plain element-wise cast;
* weight gradients: fp32 ``[C_out, kh, kw, C_in]`` -- again the arena's physical order.
"""
from __future__ import annotations

import functools
import os
from typing import Optional, Tuple

import torch

from . import require_native


@functools.lru_cache(maxsize=None)
def pick_box(out_n: int, out_h: int, out_w: int, multiple_of: int = 1, max_rows: int = 128) -> Tuple[int, int, int]:
    """Brick (bn, bh, bw) of output pixels with bn*bh*bw <= 128 that wastes the fewest MMA rows."""
    best, best_key = None, None
    total = out_n * out_h * out_w
    for bw in range(1, min(out_w, max_rows) + 1):
        for bh in range(1, min(out_h, max_rows // bw) + 1):
            bn = min(out_n, max_rows // (bw * bh))
            while bn >= 1:
                rows = bw * bh * bn
                if rows % multiple_of == 0:
                    tiles = -(-out_w // bw) * -(-out_h // bh) * -(-out_n // bn)
                    eff = total / (tiles * (max_rows if multiple_of == 1 else rows))
                    key = (round(eff, 4), bw, bh)
                    if best_key is None or key > best_key:
                        best, best_key = (bn, bh, bw), key
                    break
                bn -= 1
    if best is None:
        raise ValueError(f"no brick for output {out_n}x{out_h}x{out_w}")
    return best


def _axis_taps(out_len: int, box: int, taps: int) -> int:
    """Sum over the bricks along one axis of the filter taps whose shifted brick still overlaps the (out_len - taps + 1)
    long gradient tensor -- the data-gradient kernel skips the others (csrc/nn/conv_gemm_persistent.cu: tap_range)."""
    src = out_len - taps + 1
    total = 0
    for t0 in range(0, out_len, box):
        total += max(1, sum(1 for t in range(taps) if t0 - t + box > 0 and t0 - t < src))
    return total


@functools.lru_cache(maxsize=None)
def pick_box_dgrad(out_n: int, out_h: int, out_w: int, kh: int, kw: int) -> Tuple[int, int, int]:
    """Brick for the data gradient of a 'valid' convolution: the one with the fewest (128-row tile x filter tap) MMA
    groups once out-of-range taps are skipped.  Small images favour one-pixel bricks over many images
    (bn x 1 x 1): Toy_Net conv.10 (10x10 <- 8x8) needs 576 tile-taps per 128 images instead of 1152."""
    best, best_key = None, None
    sizes = lambda n: sorted({1, 2, 4, 8, 16, 32, 64, 128, n} & set(range(1, n + 1)))
    for bw in sizes(out_w):
        for bh in sizes(out_h):
            if bw * bh > 128:
                continue
            bn = min(out_n, 128 // (bw * bh))
            cost = -(-out_n // bn) * _axis_taps(out_h, bh, kh) * _axis_taps(out_w, bw, kw)
            key = (-cost, bw, bh)
            if best_key is None or key > best_key:
                best, best_key = (bn, bh, bw), key
    return best


def transpose_taps(w_km: torch.Tensor, taps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[C_out, taps*C_in] -> [C_in, taps*C_out] (bf16): the data gradient's K-major weight operand."""
    c_out, c_in = w_km.shape[0], w_km.shape[1] // taps
    if out is None:
        out = torch.empty((c_in, taps * c_out), dtype=w_km.dtype, device=w_km.device)
    require_native().transpose_taps(w_km, out, taps)
    return out


def use_persistent(flag: Optional[bool] = None) -> bool:
    """Persistent (one CTA per SM, double-buffered TMEM) GEMM kernels; DDPT_GEMM_PERSISTENT=0 disables."""
    if flag is not None:
        return flag
    return os.environ.get("DDPT_GEMM_PERSISTENT", "1") != "0"


@functools.lru_cache(maxsize=None)
def pick_box_halo(out_n: int, out_h: int, out_w: int, kh: int = 3, kw: int = 3,
                  max_slot_rows: int = 1 << 30) -> Tuple[Tuple[int, int, int], float]:
    """Output brick for the halo-reuse kernel: 8 pixels wide, box_h * box_n <= 16 groups, halo tile (and the
    16-group descriptor view) within ``max_slot_rows`` 128-byte rows; returns (box, fraction of useful MMA rows)."""
    best, best_key = None, None
    total = out_n * out_h * out_w
    for bh in range(1, min(out_h, 16) + 1):
        for bn in range(1, min(out_n, 16 // bh) + 1):
            halo_rows = (bh + kh - 1) * bn * 8             # one brick per filter column: row halo only
            view_rows = (16 + (kh - 1) * bn) * 8
            if max(halo_rows, view_rows) > max_slot_rows:
                continue
            tiles = -(-out_w // 8) * -(-out_h // bh) * -(-out_n // bn)
            key = (round(total / (tiles * 128), 3), bh)
            if best_key is None or key > best_key:
                best, best_key = (bn, bh, 8), key
    if best is None:
        return None, 0.0
    return best, best_key[0]


def use_cta_pair(flag: Optional[bool] = None) -> bool:
    """Run the persistent kernel as CTA pairs (tcgen05 ``cta_group::2``: two SMs share one 256-row MMA, each stages
    half of the weight tile); 128/256-wide tiles with streamed weights only.  DDPT_GEMM_PAIR=0/1 overrides."""
    if flag is not None:
        return flag
    # measured (profiles/kernel_probe_b1024.json): conv.10 forward 98 vs 111 us, conv.10 dgrad 156 vs 164 us,
    # conv.7 forward 51 vs 60 us, conv.7 dgrad 72 vs 86 us; never slower where it applies -> on by default
    return os.environ.get("DDPT_GEMM_PAIR", "1") == "1"


def _gemm_mode(persistent: Optional[bool], halo: Optional[bool], out_nhw, kh: int, kw: int, c_k: int, n_total: int,
               bn: int, pair: Optional[bool] = None):
    """(mode, box): 0 per-tile, 1 persistent, 2 persistent + halo reuse (see csrc/nn/conv_gemm_halo.cu),
    3 persistent as CTA pairs (cta_group::2)."""
    if not use_persistent(persistent):
        return 0, pick_box(*out_nhw)
    env = os.environ.get("DDPT_GEMM_HALO", "1")
    if halo is None and env == "force":
        halo = True
    want_halo = halo if halo is not None else env != "0"
    if want_halo and kh * kw > 1:
        b_bytes = kh * kw * c_k * bn * 2
        resident = n_total == bn and b_bytes <= 150 * 1024
        # resident weights leave (227 KiB - weights) / 3 per activation slot (three bricks in flight)
        max_rows = ((227 * 1024 - 2048 - b_bytes) // 3) // 128 if resident else 1 << 30
        box, eff = pick_box_halo(*out_nhw, kh, kw, max_rows)
        # measured (profiles/kernel_probe_b1024.json): the halo variant wins when the weights are resident (Toy_Net
        # conv.3 forward 78 vs 115 us with exact 8-pixel tiling; conv.3 data gradient 176 vs 188 us even with a third
        # of the MMA rows padding); streamed weights make it lose to the persistent kernel, so those stay on mode 1/3
        if box is not None and (halo or (resident and eff >= 0.6)):
            return 2, box
    return (3 if bn >= 128 and use_cta_pair(pair) else 1), pick_box(*out_nhw)


def _block_n(n_total: int, prefer: Optional[int]) -> int:
    if prefer is not None:
        return prefer
    for bn in (256, 128, 64):
        if n_total % bn == 0:
            return bn
    raise ValueError(f"N={n_total} is not a multiple of 64")


def stat_row_groups(out_n: int, out_h: int, out_w: int, c_out: int, block_n: Optional[int] = None,
                    device=None) -> int:
    """Row groups (x4 rows x 2 x C floats) a conv epilogue may write into ``stat_partials``."""
    tiles_m = 0
    for bn_, bh, bw in (pick_box(out_n, out_h, out_w), pick_box_halo(out_n, out_h, out_w, 3, 3, 300)[0],
                        pick_box_halo(out_n, out_h, out_w)[0]):
        tiles_m = max(tiles_m, -(-out_w // bw) * -(-out_h // bh) * -(-out_n // bn_))
    sms = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count
    return max(tiles_m, min(tiles_m * (c_out // _block_n(c_out, block_n)), sms))


def conv_fwd(x: torch.Tensor, w_km: torch.Tensor, kh: int, kw: int, bias: Optional[torch.Tensor] = None,
             relu: bool = False, stat_partials: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             block_n: Optional[int] = None, out_dtype=None, split_k: int = 1,
             persistent: Optional[bool] = None, halo: Optional[bool] = None, pair: Optional[bool] = None):
    """'valid' stride-1 convolution.  x [N,H,W,C], w_km [C_out, kh*kw*C] -> y [N,Ho,Wo,C_out]."""
    ext = require_native()
    n, h, w, c = x.shape
    ho, wo = h - kh + 1, w - kw + 1
    c_out = w_km.shape[0]
    if out_dtype is None:
        out_dtype = x.dtype              # bf16 or fp16 activations: the output keeps the operands' 16-bit format
    if out is None:
        out = (torch.zeros if split_k > 1 else torch.empty)((n, ho, wo, c_out), dtype=out_dtype, device=x.device)
    bn = _block_n(c_out, block_n)
    mode, box = (0, pick_box(n, ho, wo)) if split_k > 1 else _gemm_mode(persistent, halo, (n, ho, wo), kh, kw, c, c_out, bn, pair)
    tiles = ext.conv_gemm(x, w_km, out.view(n * ho * wo, c_out), [n, ho, wo], list(box), [kh, kw], 1, [0, 0], False,
                          bias, relu, stat_partials, split_k, bn, mode)
    return out, tiles


def conv_dgrad(dy: torch.Tensor, w_km: torch.Tensor, kh: int, kw: int, out: Optional[torch.Tensor] = None,
               block_n: Optional[int] = None, persistent: Optional[bool] = None, halo: Optional[bool] = None,
               pair: Optional[bool] = None, w_t: Optional[torch.Tensor] = None, kwmerge: Optional[bool] = None,
               pad: int = 0):
    """Data gradient of the stride-1 convolution ('valid' unless ``pad``).  dy [N,Ho,Wo,C_out] -> dx [N,H,W,C_in].

    ``w_t`` (optional, :func:`transpose_taps` of ``w_km``): K-major weights [C_in, taps*C_out]; with 64 input channels
    they let the data gradient run on CTA pairs (DDPT_DGRAD64_PAIR=1).

    ``kwmerge`` (default: on when it applies -- 3x3, 64 input channels, 64 / 128 output channels, no explicit flavour):
    the kernel of csrc/nn/conv_dgrad_kwmerge.cu, one 128 x 192 MMA per filter ROW instead of three 128 x 64 ones."""
    ext = require_native()
    n, ho, wo, c_out = dy.shape
    h, w = ho + kh - 1 - 2 * pad, wo + kw - 1 - 2 * pad
    c_in = w_km.shape[1] // (kh * kw)
    if out is None:
        out = torch.empty((n, h, w, c_in), dtype=dy.dtype, device=dy.device)
    can_merge = kh == 3 and kw == 3 and c_in == 64 and c_out in (64, 128)
    if kwmerge is None:
        kwmerge = (can_merge and block_n is None and persistent is None and halo is None and pair is None
                   and os.environ.get("DDPT_DGRAD_KWMERGE", "1") != "0")
    if kwmerge:
        if not can_merge:
            raise ValueError("kwmerge data gradient: 3x3 filter, 64 input channels, 64 or 128 output channels")
        ext.conv_gemm(dy, w_km, out.view(n * h * w, c_in), [n, h, w], [1, 4, 30], [3, 3], -1, [pad, pad], True,
                      None, False, None, 1, 64, 4)
        return out
    if pad:
        raise ValueError("padded data gradients exist only in the kwmerge kernel")
    bn = _block_n(c_in, block_n)
    mode, box = _gemm_mode(persistent, halo, (n, h, w), kh, kw, c_out, c_in, bn, pair)
    b, b_mn = w_km, True
    if (w_t is not None and bn == 64 and halo is not True and use_persistent(persistent) and use_cta_pair(pair)
            and os.environ.get("DDPT_DGRAD64_PAIR", "0") == "1"):
        # 64 input channels: the transposed weights are a K-major operand, which the CTA-pair kernel can split 32 + 32
        # rows (one 256 x 64 MMA per instruction).  Measured 159 us vs 159 us for the two-issuer halo kernel on conv.3
        # (every input pixel crosses L2 -> SM nine times here, 3.75 times there): opt-in.
        mode, b, b_mn = 3, w_t, False
    if mode in (1, 3) and kh * kw > 1 and os.environ.get("DDPT_DGRAD_SKIP_TAPS", "1") != "0":
        box = pick_box_dgrad(n, h, w, kh, kw)       # the persistent kernels skip filter taps that miss the brick
    ext.conv_gemm(dy, b, out.view(n * h * w, c_in), [n, h, w], list(box), [kh, kw], -1, [0, 0], b_mn,
                  None, False, None, 1, bn, mode)
    return out


@functools.lru_cache(maxsize=None)
def pick_box_multitap(out_n: int, out_h: int, out_w: int, max_rows: int = 128) -> Tuple[int, int, int]:
    """Reduction brick for the halo-reuse wgrad kernel: width a multiple of 8 (one or an even number of
    8-pixel K groups per row), rows a multiple of 16 and <= max_rows, fewest wasted rows."""
    widths = [8] + [w for w in (16, 32, 64, 128) if w <= max_rows and w <= max(out_w, 8)]
    best, best_key = None, None
    total = out_n * out_h * out_w
    for bw in widths:
        for bh in range(1, min(out_h, max_rows // bw) + 1):
            for bn in range(1, min(out_n, max_rows // (bw * bh)) + 1):
                rows = bw * bh * bn
                if rows % 16:
                    continue
                bricks = -(-out_w // bw) * -(-out_h // bh) * -(-out_n // bn)
                # >= 64-row bricks first (pipeline granularity), then fewest wasted rows, then the tallest brick
                # (smallest halo overhead for the multi-tap view)
                key = (min(rows, 64), round(total / (bricks * rows), 3), bh, rows)
                if best_key is None or key > best_key:
                    best, best_key = (bn, bh, bw), key
    if best is None:
        raise ValueError(f"no multitap brick for {out_n}x{out_h}x{out_w}")
    return best


_SLABS = {}


def _slab(device, numel: int) -> torch.Tensor:
    """Cached split-K workspace (fp32); grown on demand, never inside a CUDA-graph capture."""
    key = str(device)
    buf = _SLABS.get(key)
    if buf is None or buf.numel() < numel:
        buf = _SLABS[key] = torch.empty(numel, dtype=torch.float32, device=device)
    return buf


def use_multitap(flag: Optional[bool] = None) -> bool:
    """Halo-reuse weight-gradient kernel (csrc/nn/wgrad_gemm_multitap.cu).  DDPT_WGRAD_MULTITAP=0/1 forces it off /
    on; by default (None) the caller enables it only where it measured faster (see conv_wgrad): a filter column's
    taps merge into one wide-N MMA per k-step (N = 192 for 64-wide tiles, 2 x N = 192 for 128-wide tiles)."""
    if flag is not None:
        return flag
    env = os.environ.get("DDPT_WGRAD_MULTITAP")
    return None if env is None else env == "1"


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, kh: int, kw: int, out: Optional[torch.Tensor] = None,
               split_k: Optional[int] = None, accumulate: bool = False, block_n: Optional[int] = None,
               multitap: Optional[bool] = None, out_is_zero: bool = False):
    """dW [C_out, kh, kw, C_in] fp32 = sum over pixels dy[p, :]^T x[p + tap, :].  ``out_is_zero``: the caller already
    cleared ``out`` (e.g. one memset of the whole gradient arena), so split-K partials may be added straight in."""
    ext = require_native()
    n, ho, wo, c_out = dy.shape
    c_in = x.shape[3]
    if block_n is None and c_in % 128 == 0 and kh * kw > 1:
        block_n = 128            # measured: 128-wide accumulators beat 256 for the pixel-reduction GEMM (conv.10)
    bn = _block_n(c_in, block_n)
    mt = use_multitap(multitap)
    from . import deterministic
    det = deterministic()
    if det and kh * kw > 1 and multitap is None:
        mt = True        # split-K partials go to slabs folded in order by reduce_slabs_kernel (no red.global.add)
    if mt is None:
        # measured (profiles/kernel_probe_b1024.json): the merged wide-N kernel wins for 64-wide tiles (conv.3: 99 vs
        # 169 us) and for 128-wide tiles when the un-split grid already has >= 16 CTAs (conv.10: 124 vs 176 us); with
        # fewer CTAs the deeper split-K slab reduction eats the gain (conv.7: 95 vs 87 us)
        mt = kh * kw > 1 and (bn == 64 or (bn == 128 and (c_out // 128) * kw * (c_in // bn) >= 16))
    bricks_of = lambda b: -(-wo // b[2]) * -(-ho // b[1]) * -(-n // b[0])
    sms = torch.cuda.get_device_properties(dy.device).multi_processor_count
    slab = None
    if mt:
        box = pick_box_multitap(n, ho, wo, 128 if bn <= 128 else 64)
        max_taps = 512 // bn
        groups = kw * -(-kh // max(1, min(kh, max_taps)))      # one group per filter column (x row splits)
        ctas = (c_out // 128) * groups * (c_in // bn)
        if split_k is None:      # one resident CTA per SM (the stage ring fills shared memory): a single wave
            split_k = max(1, min(bricks_of(box), sms // ctas))
        if split_k > 1:
            slab = _slab(dy.device, split_k * c_out * kh * kw * c_in)
    else:
        box = pick_box(n, ho, wo, 16)
        ctas = (c_out // 128) * kh * kw * (c_in // bn)
        if split_k is None:      # the stage ring fills shared memory (one CTA per SM): fill one wave, never 2.07
            split_k = max(1, min(bricks_of(box), sms // ctas))
        if det:
            split_k = 1          # this kernel's split-K adds with red.global (unordered): one CTA per output tile instead
    if out is None:
        out = torch.zeros((c_out, kh, kw, c_in), dtype=torch.float32, device=dy.device)
    elif not accumulate and slab is None and split_k > 1 and not out_is_zero:
        out.zero_()
    ext.wgrad_gemm(dy, x, out.view(c_out, kh, kw, c_in), [kh, kw], list(box), split_k, accumulate, bn, mt, slab)
    return out


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
               stat_partials: Optional[torch.Tensor] = None, out_dtype=None, split_k: int = 1,
               block_n: Optional[int] = None, persistent: Optional[bool] = None, pair: Optional[bool] = None):
    """y [M, N] = x [M, K] @ w [N, K]^T (+ bias)."""
    m, k = x.shape
    y, tiles = conv_fwd(x.view(1, 1, m, k), w, 1, 1, bias, relu, stat_partials, None, block_n, out_dtype, split_k,
                        persistent, None, pair)
    return y.view(m, w.shape[0]), tiles


def linear_fwd_splitk(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      stat_partials: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                      split_k: int = 8, block_n: int = 256):
    """y [M, N] bf16 = x [M, K] @ w [N, K]^T (+ bias) for layers whose tile grid is too small to fill the GPU (Toy_Net
    dense.1: 1024 x 512 outputs = 64 tiles of 128 x 64 on 148 SMs, each streaming K = 8192 and bound by its own SM's L2
    ingest).  ``split_k`` CTAs share a 128 x ``block_n`` tile's reduction and store fp32 slabs (no atomics); one
    bandwidth-bound pass folds the slabs in order, packs bf16 and emits the BatchNorm column statistics as partial rows.
    Returns ``(y, partial_rows)``."""
    ext = require_native()
    m, k = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    slabs = _slab(x.device, split_k * m * n)[: split_k * m * n].view(split_k * m, n)
    ext.conv_gemm(x.view(1, 1, m, k), w, slabs, [1, 1, m], list(pick_box(1, 1, m)), [1, 1], 1, [0, 0], False, bias, False,
                  None, split_k, block_n, 0, True)
    rows = ext.splitk_finish(slabs, split_k, out, stat_partials)
    return out, rows


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, block_n: Optional[int] = None,
                 persistent: Optional[bool] = None, pair: Optional[bool] = None):
    """dx [M, K] = dy [M, N] @ w [N, K]."""
    m, n = dy.shape
    return conv_dgrad(dy.view(1, 1, m, n), w, 1, 1, None, block_n, persistent, None, pair).view(m, w.shape[1])


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                 split_k: Optional[int] = None, accumulate: bool = False, multitap: Optional[bool] = None,
                 out_is_zero: bool = False):
    """dw [N, K] fp32 = dy [M, N]^T @ x [M, K]."""
    m, n = dy.shape
    k = x.shape[1]
    res = conv_wgrad(dy.view(1, 1, m, n), x.view(1, 1, m, k), 1, 1,
                     None if out is None else out.view(n, 1, 1, k), split_k, accumulate, None, multitap, out_is_zero)
    return res.view(n, k)


# ---- block-scaled fp8 (MXFP8) forward path ------------------------------------------------------------------------
def quantize_mxfp8(x: torch.Tensor, data_out: Optional[torch.Tensor] = None, scale_out: Optional[torch.Tensor] = None):
    """e4m3 payload (uint8) + one UE8M0 scale per 32 elements of the last dimension (csrc/amp/quantize_mxfp8.cu)."""
    return require_native().quantize_mxfp8(x.contiguous(), data_out, scale_out)


def dequantize_mxfp8(data: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """fp32 values of an MXFP8 tensor (reference / tests)."""
    vals = data.view(torch.float8_e4m3fn).float()
    s = torch.pow(2.0, scale.float() - 127.0)
    return (vals.view(*scale.shape, 32) * s.unsqueeze(-1)).view(data.shape)


def conv_fwd_mxfp8(xq: torch.Tensor, xs: torch.Tensor, wq: torch.Tensor, ws: torch.Tensor, kh: int, kw: int,
                   bias: Optional[torch.Tensor] = None, relu: bool = False,
                   stat_partials: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """'valid' stride-1 convolution on MXFP8 operands (tcgen05.mma.kind::mxf8f6f4.block_scale), bf16 output."""
    ext = require_native()
    n, h, w, c = xq.shape
    ho, wo = h - kh + 1, w - kw + 1
    c_out = wq.shape[0]
    if out is None:
        out = torch.empty((n, ho, wo, c_out), dtype=torch.bfloat16, device=xq.device)
    box = pick_box(n, ho, wo)
    tiles = ext.conv_gemm_mxfp8(xq, xs, wq, ws, out.view(n * ho * wo, c_out), [n, ho, wo], list(box), [kh, kw], bias, relu,
                                stat_partials)
    return out, tiles


def linear_fwd_mxfp8(xq, xs, wq, ws, bias=None, relu=False, stat_partials=None, out=None):
    m, k = xq.shape
    y, tiles = conv_fwd_mxfp8(xq.view(1, 1, m, k), xs.view(1, 1, m, k // 32), wq, ws, 1, 1, bias, relu, stat_partials,
                              None if out is None else out.view(1, 1, m, wq.shape[0]))
    return y.view(m, wq.shape[0]), tiles
