"""Fused sm_100a forward/backward of Toy_Net (reference model: utils/model.py:4-34).

Dataflow (NHWC bf16 activations, fp32 statistics / parameters / gradients)::

    u8 image --conv0 (CUDA cores, K=9) + stats--> y0 --bn_finalize--> bn+relu --> a0
    a0 --tcgen05 implicit-GEMM conv (+bias, +stat partials in the epilogue)--> y1
       --bn_finalize--> bn+relu+maxpool --> a1 ... (x3) ... --> a3 [B, 8192] in torch's (c,h,w) order
    a3 --tcgen05 GEMM (+bias, +stats)--> y4 --bn_finalize--> bn+relu --> a4 --head--> logits [B,10]

    backward mirrors it: head -> BN1d bwd (reduce / finalize / apply) -> tcgen05 wgrad + dgrad
    (MN-major operands, no transposes) -> BN bwd + pool/ReLU routing -> ... -> conv0 wgrad.

Two entry points share the implementation:

* ``model(x)`` (autograd): :class:`_ToyNetFunction` returns logits, autograd supplies d(logits);
  parameter gradients come back through autograd so hooks / accumulation behave as usual.
* the step engine (``ops.step_engine``) calls :meth:`FusedToyNet.forward_impl` /
  :meth:`backward_impl` directly, with cross-entropy fused into the head kernel and parameter
  gradients written straight into the DDP gradient arena.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import require_native
from . import gemm

_EPS, _MOMENTUM = 1e-5, 0.1
# (param index of conv weight, H_in, C_in, C_out, pool) for the three tensor-core conv blocks
_TC_BLOCKS = ((4, 26, 64, 128, True), (8, 12, 128, 256, False), (12, 10, 256, 512, True))


class _Workspace:
    """Static activation / scratch buffers for one batch size (CUDA-graph friendly)."""

    def __init__(self, ext, batch: int, device, act_dtype=torch.bfloat16):
        bf, f32 = act_dtype, torch.float32
        e = lambda *s, dt=bf: torch.empty(*s, dtype=dt, device=device)
        self.batch = batch
        self.y = [e(batch, 26, 26, 64), e(batch, 24, 24, 128), e(batch, 10, 10, 256), e(batch, 8, 8, 512),
                  e(batch, 1, 1, 512)]
        self.a = [e(batch, 26, 26, 64), e(batch, 12, 12, 128), e(batch, 10, 10, 256), e(batch, 8192),
                  e(batch, 512)]
        self.coef = [torch.zeros(4, c, dtype=f32, device=device) for c in (64, 128, 256, 512, 512)]
        self.bcoef = [torch.zeros(3, c, dtype=f32, device=device) for c in (64, 128, 256, 512, 512)]
        self.logits = e(batch, 10, dt=f32)
        # statistic partial rows: conv epilogues write 4 rows (warp quadrants) per M tile
        groups = gemm.stat_row_groups
        rows = [ext.conv0_partial_rows(), 4 * groups(batch, 24, 24, 128, None, device),
                4 * max(groups(batch, 10, 10, 256, None, device), groups(batch, 10, 10, 256, 128, device)),
                4 * max(groups(batch, 8, 8, 512, None, device), groups(batch, 8, 8, 512, 128, device)),
                4 * max(groups(1, 1, batch, 512, 64, device), groups(1, 1, batch, 512, 128, device))]
        chans = (64, 128, 256, 512, 512)
        self.part = torch.zeros(max(max(r * 2 * c for r, c in zip(rows, chans)),
                                    ext.bn_bwd_partial_rows() * 2 * 512), dtype=f32, device=device)
        # backward
        self.dy = [e(batch, 26, 26, 64), e(batch, 24, 24, 128), e(batch, 10, 10, 256), e(batch, 8, 8, 512),
                   e(batch, 512)]
        self.da = [e(batch, 26, 26, 64), e(batch, 12, 12, 128), e(batch, 10, 10, 256), e(batch, 8192),
                   e(batch, 512)]
        self.da3_hwc = e(batch, 4, 4, 512)                 # da[3] (flatten order c,h,w) as channels-last
        self.saved_x: Optional[torch.Tensor] = None
        # fused BatchNorm statistics: fp64 accumulators [10][2][512] (rows 0-4 forward sum / sum of squares per layer,
        # rows 5-9 backward sum dz / sum dz*y), cleared once per training step
        self.stat = torch.zeros(10, 2 * 512, dtype=torch.float64, device=device)
        self.wt3 = e(64, 9 * 128)                          # conv.3 weights as the data gradient's K-major operand
        # block-scaled fp8 copies of the activations feeding conv.7, conv.10 and dense.1 ("O1-fp8")
        self.q = {}

    def acc(self, row: int, channels: int) -> torch.Tensor:
        return self.stat[row, : 2 * channels]

    def fp8_buffers(self, k: int, like: torch.Tensor):
        if k not in self.q:
            u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=like.device)
            shape = list(like.shape)
            self.q[k] = (u8(*shape), u8(*shape[:-1], shape[-1] // 32))
        return self.q[k]


class FusedToyNet:
    def __init__(self, model, compute: str = "bf16"):
        self.ext = require_native()
        self.model = model
        self.compute = compute
        # activation / operand format: bf16 ("O1", "O1-fp8") or fp16 ("O1-fp16": Apex O1's precision in the reference,
        # utils/train.py:58) -- the same kernels; csrc/nn/toynet_elementwise.cu is built for both types and the GEMMs take
        # the format from the tensors' dtype
        self.act_dtype = torch.float16 if compute == "fp16" else torch.bfloat16
        self.params: List[torch.nn.Parameter] = list(model.parameters())
        if len(self.params) != 22:
            raise ValueError("FusedToyNet expects the 22-parameter Toy_Net")
        self.device = self.params[0].device
        self.bn = [model.conv[1], model.conv[4], model.conv[8], model.conv[11], model.dense[2]]
        self._ws: Dict[int, _Workspace] = {}
        self._shadow_cache: Dict[int, tuple] = {}
        self._w_fp8: Dict[int, tuple] = {}
        self._staging: Optional[torch.Tensor] = None
        self.launches = 0
        import os
        # BatchNorm statistics travel as partial rows folded by *_finalize kernels (ordered, bit-stable).  The alternative
        # (DDPT_FUSED_BN_STATS=1) lets the producing kernels' epilogues add into fp64 accumulators and the consuming pass
        # derive the coefficients itself -- ten launches fewer.  Measured on the same box (200 steps, graph + PDL):
        # 1.485 ms fused vs 1.472 ms with the finalize kernels: under programmatic dependent launch a finalize kernel
        # costs ~4 us, about what the contended red.global.add.f64 tail of 148 CTAs plus the consumers' prologue costs.
        # The fused path therefore stays opt-in (it is also unordered, i.e. not bit-reproducible).
        from . import deterministic
        # DDPT_DETERMINISTIC=1: every cross-block float reduction is an ordered fold of per-block rows (head, conv.0
        # weight gradient) or of split-K slabs (tensor-core weight gradients) instead of red.global; with the partial-row
        # BatchNorm statistics and the in-order block folds this makes the whole step bit-reproducible run to run
        # (reference: utils/customized.py:12-13 sets cudnn.deterministic).
        self.deterministic = deterministic()
        self._det_ws: Optional[torch.Tensor] = None
        self.dgrad64_pair = os.environ.get("DDPT_DGRAD64_PAIR", "0") == "1"
        self.fuse_conv0_bn = os.environ.get("DDPT_FUSE_CONV0_BN", "1") != "0"
        # tensor-core layers leave their bias to the BatchNorm statistics kernels (see _folds_bias); DDPT_FOLD_CONV_BIAS=0
        # keeps the add in the GEMM epilogues
        self.fold_bias = os.environ.get("DDPT_FOLD_CONV_BIAS", "1") != "0"
        self.dense_splitk = int(os.environ.get("DDPT_DENSE_SPLITK", "4"))   # measured: 128 x 128 tiles x 4 splits 27.5 us, 64-wide tiles unsplit 35.8 us
        self.fused_stats = (compute != "fp8" and not deterministic()
                            and os.environ.get("DDPT_FUSED_BN_STATS", "0") == "1")

    def count_launches(self, batch: int) -> int:
        """Kernels of ours in one training step (forward + backward + unscale + scale update + optimizer)."""
        fwd = 1 + 1 + 1 + 3 * 3 + 3              # conv0, finalize, apply | 3x(conv, finalize, apply) | dense.1 trio (head: in bwd)
        bwd = 1 + 3 + 2 + 3 * (3 + 2) + 3 + 1    # head | bn1d x3 | dense.1 wgrad+dgrad | 3x(bn x3, wgrad, dgrad) | bn0 x3 | conv0 wgrad
        bwd += 1 + 2                             # flatten -> channels-last transpose | split-K slab reductions (conv.3, conv.10 wgrad)
        if self.fuse_conv0_bn and not self.fused_stats:
            bwd -= 1                             # block 0's bn_bwd_apply lives in conv0_wgrad
        if self.dense_splitk > 1 and not self.fused_stats and batch >= 256 and self.compute != "fp8":
            fwd += 1                             # dense.1: split-K GEMM + slab fold / statistics pass
        if self.fused_stats:
            fwd -= 5                             # no bn_finalize launches
            bwd -= 5                             # no bn_bwd_finalize launches
        return fwd + bwd + 3                     # + unscale, scale update, optimizer (profiles/launches_step_b1024_v9.txt)

    # -- helpers ------------------------------------------------------------------------------------
    def det_partials(self) -> Optional[torch.Tensor]:
        """Partial-row workspace of deterministic mode (None in fast mode)."""
        if not self.deterministic:
            return None
        if self._det_ws is None:
            self._det_ws = torch.zeros(self.ext.det_workspace_floats(), dtype=torch.float32, device=self.device)
        return self._det_ws

    def workspace(self, batch: int) -> _Workspace:
        ws = self._ws.get(batch)
        if ws is None:
            ws = self._ws[batch] = _Workspace(self.ext, batch, self.device, self.act_dtype)
        return ws

    def _weight_bf16(self, pidx: int) -> torch.Tensor:
        """bf16 [C_out, kh*kw*C_in] (channels-last physical order) or [N, K] for linear weights."""
        p = self.params[pidx]
        fused_opt = getattr(self.model, "_ddpt_fused_optim", None)
        if fused_opt is not None and fused_opt.arena.channels_last and fused_opt.shadow.dtype == self.act_dtype:
            return fused_opt.shadow_view(pidx).view(p.shape[0], -1)
        key = (p.data_ptr(), p._version)
        hit = self._shadow_cache.get(pidx)
        if hit is not None and hit[0] == key:
            return hit[1]
        src = p.detach()
        if src.dim() == 4:
            src = src.permute(0, 2, 3, 1)
        w = src.reshape(p.shape[0], -1).to(self.act_dtype).contiguous()
        self._shadow_cache[pidx] = (key, w)
        return w

    def _weight_fp8(self, pidx: int):
        """MXFP8 copy (e4m3 + UE8M0 per 32 along K) of a weight matrix, re-quantised from the bf16 shadow."""
        w = self._weight_bf16(pidx)
        if pidx not in self._w_fp8:
            self._w_fp8[pidx] = (torch.empty(w.shape, dtype=torch.uint8, device=w.device),
                                 torch.empty((w.shape[0], w.shape[1] // 32), dtype=torch.uint8, device=w.device))
        q, sc = self._w_fp8[pidx]
        self.ext.quantize_mxfp8(w, q, sc)
        return q, sc

    def _grad_view(self, flat: torch.Tensor, offset: int, p: torch.Tensor) -> torch.Tensor:
        v = flat[offset: offset + p.numel()]
        if p.dim() == 4:
            co, ci, kh, kw = p.shape
            return v.view(co, kh, kw, ci)            # physical order of the kernels' output
        return v.view(p.shape)

    def staging_grads(self) -> List[torch.Tensor]:
        """Per-parameter fp32 buffers (conv weights as [C_out,kh,kw,C_in]) inside one flat allocation."""
        if self._staging is None:
            total, offs = 0, []
            for p in self.params:
                offs.append(total)
                total += (p.numel() + 3) // 4 * 4
            self._staging = torch.zeros(total, dtype=torch.float32, device=self.device)
            self._staging_views = [self._grad_view(self._staging, o, p) for o, p in zip(offs, self.params)]
        return self._staging_views

    def arena_grads(self, arena) -> List[torch.Tensor]:
        if not arena.channels_last:
            raise RuntimeError("fused Toy_Net writes channels-last weight gradients; arena must match")
        return [self._grad_view(arena.grad_arena, arena.offset_of[i], p) for i, p in enumerate(self.params)]

    # -- forward ---------------------------------------------------------------------------------------
    def _bn_coeffs(self, k: int, ws: _Workspace, train: bool, rows: int, count: float,
                   conv_bias: Optional[torch.Tensor] = None) -> None:
        """``conv_bias``: the producing GEMM left its bias out (see :meth:`_folds_bias`); the statistics kernels put it back
        where it is observable -- running_mean in training, the shift of the running statistics in eval."""
        bn = self.bn[k]
        if train:
            self.ext.bn_finalize(ws.part, rows, count, bn.weight, bn.bias, _EPS, _MOMENTUM, ws.coef[k],
                                 bn.running_mean, bn.running_var, bn.num_batches_tracked, conv_bias)
        else:
            self.ext.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, _EPS, ws.coef[k], conv_bias)

    def _folds_bias(self, fused: bool, fp8_layer: bool) -> bool:
        """A bias in front of a BatchNorm is a per-channel constant the normalisation removes again: the output, the ReLU
        mask and every gradient are those of the bias-free convolution (and d loss / d bias is exactly zero).  The
        tensor-core layers therefore skip the add -- 64 loads + 64 adds per thread and tile in an epilogue that paces
        conv.3's forward (profiles/ncu_r2_fwd3_halo_step.txt) -- store bias-free activations, and hand the bias to the
        statistics kernels.  Not with the fp64-accumulator statistics (opt-in) or the MXFP8 layers, whose passes do not
        take the bias."""
        return self.fold_bias and not fused and not fp8_layer

    def _bn_act(self, k: int, ws: _Workspace, train: bool, fused: bool, rows: int, count: float, pool: bool,
                flatten: bool, q, conv_bias: Optional[torch.Tensor] = None) -> None:
        """BatchNorm (batch or running statistics) + ReLU (+ 2x2 max-pool, + flatten) of block k: y[k] -> a[k]."""
        ext, bn = self.ext, self.bn[k]
        if fused:      # the pass derives scale/shift from the accumulators itself and publishes coef / running stats
            ext.bn_relu_pool(ws.y[k], ws.coef[k], ws.a[k], pool, flatten, q[0], q[1], ws.acc(k, bn.num_features), count,
                             bn.weight, bn.bias, _EPS, _MOMENTUM, bn.running_mean, bn.running_var,
                             bn.num_batches_tracked)
        else:
            self._bn_coeffs(k, ws, train, rows, count, conv_bias)
            ext.bn_relu_pool(ws.y[k], ws.coef[k], ws.a[k], pool, flatten, q[0], q[1])

    def forward_impl(self, x: torch.Tensor, train: bool, target: Optional[torch.Tensor] = None,
                     metrics: Optional[torch.Tensor] = None, head: bool = True) -> torch.Tensor:
        ext, P = self.ext, self.params
        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError("Toy_Net takes single-channel 28x28 images")
        x = x.contiguous()
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        batch = x.numel() // 784
        ws = self.workspace(batch)
        ws.saved_x = x
        part = ws.part if train else None
        fused = train and self.fused_stats
        if fused:
            ws.stat.zero_()                                    # all ten accumulator rows (forward and backward)
        stat_out = (lambda k, c: ws.acc(k, c)) if fused else (lambda k, c: part)
        # block 0: direct conv + BN + ReLU
        rows = ext.conv0_fwd(x, P[0].detach().reshape(64, 9).contiguous(), P[1].detach(), ws.y[0], stat_out(0, 64))
        self._bn_act(0, ws, train, fused, rows, batch * 676.0, False, False, (None, None))
        # blocks 1..3: tcgen05 implicit GEMM + BN + ReLU (+pool)
        act = ws.a[0]
        for k, (pidx, h_in, c_in, c_out, pool) in enumerate(_TC_BLOCKS, start=1):
            h_out = h_in - 2
            bias = P[pidx + 1].detach()
            fold = self._folds_bias(fused, self.compute == "fp8" and c_in % 128 == 0)
            if self.compute == "fp8" and c_in % 128 == 0:
                xq, xs = ws.fp8_buffers(k, act)            # written by the previous block's fused BN+ReLU(+pool) pass
                wq, wsc = self._weight_fp8(pidx)
                _, tiles = gemm.conv_fwd_mxfp8(xq, xs, wq, wsc, 3, 3, P[pidx + 1].detach(), False, part, ws.y[k])
            else:
                _, tiles = gemm.conv_fwd(act, self._weight_bf16(pidx), 3, 3, None if fold else bias, False,
                                         stat_out(k, c_out), ws.y[k])
            q = ws.fp8_buffers(k + 1, ws.a[k]) if self.compute == "fp8" else (None, None)   # consumer: conv.7 / conv.10 / dense.1
            self._bn_act(k, ws, train, fused, 4 * tiles, batch * float(h_out * h_out), pool, k == 3, q,
                         bias if fold else None)
            act = ws.a[k]
        # dense.1 + BN1d + ReLU
        y4 = ws.y[4].view(batch, 512)
        bias4 = P[17].detach()
        fold4 = self._folds_bias(fused, self.compute == "fp8")
        if self.compute == "fp8":
            xq, xs = ws.fp8_buffers(4, ws.a[3])
            wq, wsc = self._weight_fp8(16)
            _, tiles = gemm.linear_fwd_mxfp8(xq, xs, wq, wsc, P[17].detach(), False, part, y4)
            rows4 = 4 * tiles
        elif self.dense_splitk > 1 and not fused and batch >= 256:
            # 8 x 8 output tiles cannot fill 148 SMs and each streams K = 8192 through one SM's L2 port: split the
            # reduction over `dense_splitk` CTAs per 128 x 128 tile (fp32 slabs), fold + pack + statistics in one pass
            _, rows4 = gemm.linear_fwd_splitk(ws.a[3], self._weight_bf16(16), None if fold4 else bias4, part, y4,
                                              split_k=self.dense_splitk, block_n=128)
        else:
            _, tiles = gemm.conv_fwd(ws.a[3].view(1, 1, batch, 8192), self._weight_bf16(16), 1, 1,
                                     None if fold4 else bias4, False, stat_out(4, 512), y4.view(1, 1, batch, 512),
                                     block_n=64)
            rows4 = 4 * tiles
        self._bn_act(4, ws, train, fused, rows4, float(batch), False, False, (None, None), bias4 if fold4 else None)
        if not head:       # training step engine: the backward head launch also produces loss / accuracy
            return ws.logits
        # head: logits (+ fused cross-entropy metrics when a target is given and no gradient is needed here)
        ext.head(ws.a[4], P[20].detach(), P[21].detach(), target, ws.logits, metrics, None, None, None, None, None,
                 float(batch), self.det_partials() if metrics is not None else None)
        return ws.logits

    # -- backward --------------------------------------------------------------------------------------
    def backward_impl(self, batch: int, grads: List[torch.Tensor], dlogits: Optional[torch.Tensor] = None,
                      target: Optional[torch.Tensor] = None, loss_scale: Optional[torch.Tensor] = None,
                      on_ready: Optional[Callable[[int], None]] = None,
                      metrics: Optional[torch.Tensor] = None, side_stream=None) -> None:
        """Write d(loss)/d(param) into ``grads`` (zero-initialised by the caller).

        Either ``dlogits`` (autograd path) or ``target`` (fused CE with the reference's extra 1/B and the
        device-resident ``loss_scale``) selects the loss gradient.  ``on_ready(param_index)`` fires as
        soon as a parameter's gradient is final -- in the order DDP buckets expect."""
        ext, P, ws = self.ext, self.params, self.workspace(batch)
        ready = on_ready or (lambda i: None)
        x = ws.saved_x
        main = torch.cuda.current_stream()

        def off_critical_path(fn):
            """Weight gradients are leaves of the backward graph: with a side stream they overlap the
            bandwidth-bound BatchNorm passes and the data-gradient GEMM of the next layer (their CTAs need few
            registers, so they share SMs with the element-wise kernels)."""
            if side_stream is None:
                fn()
                return
            ev = torch.cuda.Event()
            ev.record(main)
            side_stream.wait_event(ev)
            with torch.cuda.stream(side_stream):
                fn()

        ext.head(ws.a[4], P[20].detach(), P[21].detach(), target, None, metrics, loss_scale, ws.da[4], grads[20],
                 grads[21], dlogits, float(batch), self.det_partials())
        ready(21); ready(20)
        # BN1d backward -> dy4
        # biases in front of a BatchNorm (dense.1, conv.3/7/10) have an identically zero gradient: their slots keep
        # the zero the gradient buffers are cleared to, and the apply pass skips the per-channel sum (dbias_idx=None)
        self._bn_backward(4, ws, ws.da[4], grads, False, False, float(batch), dbias_idx=None)
        ready(19); ready(18); ready(17)
        dy4 = ws.dy[4]
        off_critical_path(lambda: gemm.linear_wgrad(dy4, ws.a[3], out=grads[16], out_is_zero=True))
        ready(16)
        gemm.conv_dgrad(dy4.view(1, 1, batch, 512), self._weight_bf16(16), 1, 1, ws.da[3].view(1, 1, batch, 8192))
        # the Linear's data gradient is in flatten order (c, h, w); one 16 MB transpose (L2-resident input) lets the
        # pooled BatchNorm passes read 16-byte channel vectors instead of gathering 2-byte elements (113 -> ~50 us)
        ext.chw_to_hwc(ws.da[3], ws.da3_hwc, 512, 16)
        g = ws.da3_hwc
        for k in (3, 2, 1):
            pidx, h_in, c_in, c_out, pool = _TC_BLOCKS[k - 1]
            h_out = h_in - 2
            self._bn_backward(k, ws, g, grads, pool, False, batch * float(h_out * h_out), dbias_idx=None)
            ready(pidx + 3); ready(pidx + 2); ready(pidx + 1)
            off_critical_path(lambda k=k, pidx=pidx: gemm.conv_wgrad(ws.dy[k], ws.a[k - 1], 3, 3, out=grads[pidx],
                                                                     out_is_zero=True))
            ready(pidx)
            w_t = None
            if c_in == 64 and self.dgrad64_pair:     # conv.3 on CTA pairs needs the K-major (tap-transposed) weights
                w_t = gemm.transpose_taps(self._weight_bf16(pidx), 9, ws.wt3)
            g = gemm.conv_dgrad(ws.dy[k], self._weight_bf16(pidx), 3, 3, ws.da[k - 1], w_t=w_t)
        if self.fuse_conv0_bn and not self.fused_stats:
            # block 0: conv.0 has no data gradient, so dy0 would be written by the BatchNorm apply pass only for
            # conv.0's weight gradient to read back.  Reduce + finalize as usual; the apply folds into conv0_wgrad.
            bn0 = self.bn[0]
            rows = ext.bn_bwd_reduce(ws.y[0], g, ws.coef[0], ws.part, False, False)
            ext.bn_bwd_finalize(ws.part, rows, batch * 676.0, bn0.weight, ws.coef[0], grads[2], grads[3], ws.bcoef[0], False)
            ready(3); ready(2)
            ext.conv0_wgrad(x, ws.dy[0], grads[0].view(-1), grads[1], self.det_partials(), ws.y[0], g, ws.coef[0],
                            ws.bcoef[0])
        else:
            self._bn_backward(0, ws, g, grads, False, False, batch * 676.0, dbias_idx=None)
            ready(3); ready(2)
            ext.conv0_wgrad(x, ws.dy[0], grads[0].view(-1), grads[1], self.det_partials())
        if side_stream is not None:                       # join: everything after backward sees all weight gradients
            ev = torch.cuda.Event()
            ev.record(side_stream)
            main.wait_event(ev)
        ready(1); ready(0)

    def _bn_backward(self, k: int, ws: _Workspace, g: torch.Tensor, grads, pool: bool, flatten: bool, count: float,
                     dbias_idx: Optional[int]) -> None:
        ext, bn = self.ext, self.bn[k]
        gamma_idx = {0: 2, 1: 6, 2: 10, 3: 14, 4: 18}[k]
        dy = ws.dy[k] if k < 4 else ws.dy[4].view(ws.batch, 1, 1, 512)
        dbias = grads[dbias_idx] if dbias_idx is not None else None
        if self.fused_stats:
            acc = ws.acc(5 + k, bn.num_features)             # cleared with the forward accumulators at the start of the step
            ext.bn_bwd_reduce(ws.y[k], g, ws.coef[k], acc, pool, flatten)
            ext.bn_bwd_apply(ws.y[k], g, ws.coef[k], ws.bcoef[k], dy, dbias, pool, flatten, acc, count, bn.weight,
                             grads[gamma_idx], grads[gamma_idx + 1])
            return
        rows = ext.bn_bwd_reduce(ws.y[k], g, ws.coef[k], ws.part, pool, flatten)
        ext.bn_bwd_finalize(ws.part, rows, count, bn.weight, ws.coef[k], grads[gamma_idx], grads[gamma_idx + 1],
                            ws.bcoef[k], False)
        ext.bn_bwd_apply(ws.y[k], g, ws.coef[k], ws.bcoef[k], dy, dbias, pool, flatten)

    # -- autograd entry ------------------------------------------------------------------------------------
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        train = self.model.training
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.params)
        if not needs_grad:
            return self.forward_impl(x, train).clone()
        return _ToyNetFunction.apply(x, self, *self.params)


class _ToyNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fused: FusedToyNet, *params):
        ctx.fused = fused
        ctx.batch = x.numel() // 784
        logits = fused.forward_impl(x, fused.model.training)
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        fused: FusedToyNet = ctx.fused
        grads = fused.staging_grads()
        fused._staging.zero_()
        fused.backward_impl(ctx.batch, grads, dlogits=dlogits.contiguous().float())
        out = []
        for p, g in zip(fused.params, grads):
            out.append(g.permute(0, 3, 1, 2) if p.dim() == 4 else g)
        # fresh tensors every call so autograd may keep them (``p.grad = g`` when grad was None)
        return (None, None, *[o.clone(memory_format=torch.preserve_format) for o in out])


def attach(model, compute: str = "bf16") -> bool:
    """Install the fused path on a Toy_Net living on a CUDA device."""
    if not next(model.parameters()).is_cuda:
        return False
    model._fused = FusedToyNet(model, compute)
    return True
