"""Autograd wrappers of the U-Net glue kernels: GroupNorm (+ time-embedding add, + SiLU) on
channels-last bf16 activations and GEGLU.  The normalisation's affine parameters belong to the frozen U-Net
(train/ppft_train.py:569-578), so the backward produces the input gradient only.

Reference op sequences: scripts/lib/original_unet.py:440-453 (ResnetBlock2D.forward), :826 (Transformer2DModel.norm),
:1416 (conv_norm_out), :708-729 (GEGLU).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import AqualoraError


class _GroupNormNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, add_bc, groups, eps, silu):
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or (add_bc is not None and ctx.needs_input_grad[3]):
            raise AqualoraError("group_norm_nhwc: gamma / beta / the added embedding must be frozen (no gradient is produced for them)")
        x = x.contiguous(memory_format=torch.channels_last)          # no-op inside the channels_last U-Net
        add_bc = None if add_bc is None else add_bc.contiguous()
        y, stats = ops.group_norm_nhwc_fwd(x, gamma, beta, groups, eps, silu, add_bc)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(x, gamma, beta, stats, *([] if add_bc is None else [add_bc]))
        ctx.cfg = (groups, eps, silu, add_bc is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        groups, eps, silu, has_add = ctx.cfg
        saved = ctx.saved_tensors
        x, gamma, beta, stats = saved[:4]
        add_bc = saved[4] if has_add else None
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = ops.group_norm_nhwc_bwd(dy, x, gamma, beta, stats, groups, eps, silu, add_bc)
        return dx, None, None, None, None, None, None


def group_norm_nhwc(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool = False,
                    add_bc: torch.Tensor | None = None) -> torch.Tensor:
    """[silu](GroupNorm(x + add_bc[:, :, None, None])) for x [B, C, H, W] bf16 (kept / made channels_last)."""
    return _GroupNormNHWC.apply(x, gamma, beta, add_bc, int(groups), float(eps), bool(silu))


class _Geglu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj):
        if proj.stride(-1) != 1:
            proj = proj.contiguous()
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(proj)
        return ops.geglu_fwd(proj)

    @staticmethod
    def backward(ctx, g_out):
        (proj,) = ctx.saved_tensors
        return ops.geglu_bwd(proj, g_out.contiguous())


def geglu(proj: torch.Tensor) -> torch.Tensor:
    """proj[..., :F] * gelu(proj[..., F:]) (erf gelu) in one pass, one pass for the backward."""
    return _Geglu.apply(proj)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise AqualoraError("layer_norm: gamma / beta must be frozen (no gradient is produced for them)")
        x = x.contiguous()
        y, stats = ops.layer_norm_fwd(x, gamma, beta, eps, save_stats=ctx.needs_input_grad[0])
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(x, gamma, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, stats = ctx.saved_tensors
        return ops.layer_norm_bwd(dy.contiguous(), x, gamma, stats), None, None, None


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    """LayerNorm over the last dimension of bf16 token rows (BasicTransformerBlock.norm1/2/3), frozen affine parameters."""
    return _LayerNorm.apply(x, gamma, beta, float(eps))


class _AddBias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, bias):
        if ctx.needs_input_grad[2]:
            raise AqualoraError("residual_add_bias: the bias must be frozen")
        cl = torch.channels_last
        return ops.add_bias_nhwc(a.contiguous(memory_format=cl), b.contiguous(memory_format=cl), bias)

    @staticmethod
    def backward(ctx, g):
        return (g if ctx.needs_input_grad[0] else None), (g if ctx.needs_input_grad[1] else None), None


def residual_add_bias(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """a + b + bias[None, :, None, None] in one pass (ResnetBlock2D's closing add with the convolution biases folded in)."""
    return _AddBias.apply(a, b, bias)
