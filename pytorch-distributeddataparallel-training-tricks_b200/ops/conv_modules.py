"""Tensor-core 1x1 convolutions for arbitrary models (``nn.Conv2d`` drop-in): the framework's own GEMM kernels
behind autograd.

The reference's DDP / AMP / Lookahead stack wraps any ``nn.Module`` (utils/train.py:58-61); the hand-written
Toy_Net path is one model.  For everything else -- ResNet-50 is BASELINE.json config 4 -- the pointwise
convolutions (36 of ResNet-50's 53 convolutions, ~51 % of its FLOPs) are plain GEMMs over the NHWC pixel matrix, and
run here on the same tcgen05 persistent / CTA-pair kernels as Toy_Net's layers:

    forward   y[M, C_out] = x[M, C_in] @ W^T            ops.gemm.linear_fwd   (TMA + tcgen05, bf16 in / bf16 out)
    dgrad     dx[M, C_in] = dy[M, C_out] @ W            ops.gemm.linear_dgrad (MN-major weights, no transpose)
    wgrad     dW[C_out, C_in] = dy^T @ x                ops.gemm.linear_wgrad (both operands MN-major, fp32 out)

``M = N*H*W``; a channels-last activation *is* that matrix, so there is no layout change on either side.  The bf16
weights come from the fused optimizer's shadow arena when one is attached (written by the optimizer kernel itself),
else from a per-step cast.  Stride-2 pointwise convolutions (ResNet shortcut projections) subsample first.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import gemm, native_available


class _PointwiseConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d: torch.Tensor, weight: torch.Tensor, w_bf16: torch.Tensor, bias: Optional[torch.Tensor]):
        y, _ = gemm.linear_fwd(x2d, w_bf16, bias)
        ctx.save_for_backward(x2d, w_bf16)
        ctx.has_bias = bias is not None
        ctx.w_shape = weight.shape
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        x2d, w_bf16 = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dx = gemm.linear_dgrad(dy, w_bf16) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            c_out, c_in = dy.shape[1], x2d.shape[1]
            if c_out % 128 == 0:                      # the wgrad kernel tiles its M (= C_out) dimension by 128
                dw = gemm.linear_wgrad(dy, x2d)
            elif c_in % 128 == 0:                     # dW^T = x^T dy has M = C_in: same kernel, operands swapped
                dw = gemm.linear_wgrad(x2d, dy).t().contiguous()
            else:                                     # 64 -> 64 (ResNet-50 has one such layer): library GEMM
                dw = (dy.t() @ x2d).float()
            dw = dw.view(ctx.w_shape)
        db = dy.float().sum(0) if ctx.has_bias and ctx.needs_input_grad[3] else None
        return dx, dw, None, db


class TensorCoreConv1x1(nn.Conv2d):
    """``nn.Conv2d(c_in, c_out, 1)`` whose CUDA bf16 path is the framework's tcgen05 GEMM.

    Same parameters, same ``state_dict``, same initialisation as ``nn.Conv2d``; any existing ``nn.Conv2d`` with a 1x1
    kernel can be switched in place with :func:`use_tensor_core_pointwise_convs`.  Falls back to the stock
    implementation whenever the fused path does not apply (CPU, fp32 without autocast, channel counts that are not
    multiples of 64, groups, dilation, padding)."""

    _ddpt_root: Optional[weakref.ReferenceType] = None
    _ddpt_pidx: int = -1
    fused_calls = 0

    def _eligible(self, x: torch.Tensor) -> bool:
        return (x.is_cuda and x.dim() == 4 and self.kernel_size == (1, 1) and self.groups == 1
                and self.padding in ((0, 0), 0) and self.dilation == (1, 1) and self.stride[0] == self.stride[1]
                and self.stride[0] in (1, 2) and self.in_channels % 64 == 0 and self.out_channels % 64 == 0
                and (x.dtype == torch.bfloat16 or (torch.is_autocast_enabled() and
                                                   torch.get_autocast_dtype("cuda") == torch.bfloat16))
                and native_available())

    def _weight_bf16(self) -> torch.Tensor:
        root = self._ddpt_root() if self._ddpt_root is not None else None
        fused_opt = getattr(root, "_ddpt_fused_optim", None) if root is not None else None
        if fused_opt is not None and self._ddpt_pidx >= 0 and fused_opt.params[self._ddpt_pidx] is self.weight:
            return fused_opt.shadow_view(self._ddpt_pidx).view(self.out_channels, self.in_channels)
        return self.weight.detach().reshape(self.out_channels, self.in_channels).to(torch.bfloat16)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self._eligible(x):
            return super().forward(x)
        if self.stride[0] == 2:
            x = x[:, :, ::2, ::2]
        n, c, h, w = x.shape
        xn = x.permute(0, 2, 3, 1)                       # channels-last storage: this view is already contiguous
        x2d = xn.reshape(n * h * w, c)
        if x2d.dtype != torch.bfloat16:
            x2d = x2d.to(torch.bfloat16)
        y2d = _PointwiseConvFn.apply(x2d.contiguous(), self.weight, self._weight_bf16(), self.bias)
        TensorCoreConv1x1.fused_calls += 1
        return y2d.view(n, h, w, self.out_channels).permute(0, 3, 1, 2)       # NCHW view, channels-last strides


def use_tensor_core_pointwise_convs(model: nn.Module) -> int:
    """Switch every eligible 1x1 ``nn.Conv2d`` of ``model`` to :class:`TensorCoreConv1x1` in place (parameters, buffers
    and ``state_dict`` keys are untouched) and remember each layer's parameter index so the bf16 weights can be read
    from the fused optimizer's shadow arena.  Returns the number of layers switched."""
    index = {id(p): i for i, p in enumerate(model.parameters())}
    root = weakref.ref(model)
    switched = 0
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1) and m.groups == 1 and \
                m.in_channels % 64 == 0 and m.out_channels % 64 == 0:
            if not isinstance(m, TensorCoreConv1x1):
                m.__class__ = TensorCoreConv1x1
            m._ddpt_root, m._ddpt_pidx = root, index.get(id(m.weight), -1)
            switched += 1
    return switched


def pointwise_flop_share(model: nn.Module, input_hw=(224, 224), in_channels: int = 3) -> float:
    """Fraction of the model's convolution + linear multiply-accumulates that run on TensorCoreConv1x1 layers
    (forward; backward doubles both sides alike), measured with one batch-1 forward pass."""
    macs = {"own": 0.0, "all": 0.0}
    hooks = []

    def hook(mod, inp, out):
        if isinstance(mod, nn.Conv2d):
            k = mod.kernel_size[0] * mod.kernel_size[1] * mod.in_channels // mod.groups
            m = out.numel() * k
            macs["all"] += m
            if isinstance(mod, TensorCoreConv1x1) and mod.in_channels % 64 == 0 and mod.out_channels % 64 == 0:
                macs["own"] += m
        elif isinstance(mod, nn.Linear):
            macs["all"] += out.numel() * mod.in_features

    for mod in model.modules():
        if isinstance(mod, (nn.Conv2d, nn.Linear)):
            hooks.append(mod.register_forward_hook(hook))
    with torch.no_grad():
        was_training = model.training
        try:
            dev = next(model.parameters()).device
            model.eval()
            x = torch.zeros(1, in_channels, *input_hw, device=dev)
            if dev.type == "cuda":
                x = x.contiguous(memory_format=torch.channels_last)
            model(x)
        finally:
            model.train(was_training)
            for h in hooks:
                h.remove()
    return macs["own"] / max(macs["all"], 1.0)
