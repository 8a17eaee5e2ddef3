"""Losses of the latent-watermark pretraining step (train/latent_wm_pretrain.py) on the CUDA kernels of csrc/losses.cu.

    PRVL_loss(img1, img2)                         :42-50   max over positions of the 32 x 32 box mean of mean_c |img1 - img2|
    binary_cross_entropy_with_logits(x, target)   :200     mean reduction (the message loss)
Both are differentiable; CPU tensors raise (no PyTorch fallback).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import AqualoraError


class _PrvlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2):
        loss, state = ops.prvl_loss_fwd(img1, img2)
        ctx.save_for_backward(img1, img2, state)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        img1, img2, state = ctx.saved_tensors
        g1, g2 = ops.prvl_loss_bwd(img1, img2, state, g, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g1, g2


def PRVL_loss(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    if not img1.is_cuda:
        raise AqualoraError("PRVL_loss: CPU tensor passed; aqualora_b200 has no CPU fallback")
    return _PrvlFn.apply(img1.float().contiguous(), img2.float().contiguous())


class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets):
        loss, g = ops.bce_logits(logits, targets, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(g)
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return (g * gl if g is not None else None), None


def binary_cross_entropy_with_logits(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    if not logits.is_cuda:
        raise AqualoraError("binary_cross_entropy_with_logits: CPU tensor passed; aqualora_b200 has no CPU fallback")
    return _BceFn.apply(logits.float(), targets.float())
