"""Drop-in counterparts of the reference's `utils/models.py` classes (same constructor signatures, attribute names and
state-dict keys), backed by the sm_100a kernels.  CPU tensors raise: there is no PyTorch fallback.

    MapperNet(input_size, output_size, std=1.)          utils/models.py:98-115
    SecretEncoder(secret_len, base_res=32, resolution=64)  utils/models.py:51-81   (differentiable: trained in
                                                           train/latent_wm_pretrain.py, no_grad in train/ppft_train.py:994-996)
    SecretDecoder(output_size)                          utils/models.py:84-96   -> aqualora_b200/decoder.py
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.init as init

from . import ops
from ._lib import AqualoraError


class _MapperFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msg, emb):
        ctx.save_for_backward(msg)
        return ops.mapper_fwd(msg, emb, round_bf16=False)

    @staticmethod
    def backward(ctx, g):
        (msg,) = ctx.saved_tensors
        g_emb = torch.zeros((msg.shape[1], g.shape[1]), dtype=torch.float32, device=g.device)
        ops.mapper_bwd(msg, g.contiguous().float(), g_emb)
        return None, g_emb


class MapperNet(nn.Module):
    """scale = 1 + sum_i msg[:, i] * E[i, :] / sqrt(bits); E orthogonal-init, rows normalised to unit std."""

    def __init__(self, input_size=16, output_size=64, std=1.0):
        super().__init__()
        self.input_size = input_size
        self.output_size = output_size
        self.bit_embeddings = nn.Embedding(input_size, output_size)
        init.orthogonal_(self.bit_embeddings.weight)
        self.bit_embeddings.weight.data = self.bit_embeddings.weight.data / self.bit_embeddings.weight.data.std(dim=1, keepdim=True)
        self.bit_embeddings.weight.data = self.bit_embeddings.weight.data * std

    def forward(self, x):
        if not x.is_cuda:
            raise AqualoraError("MapperNet: CPU tensor passed; aqualora_b200 has no CPU fallback")
        return _MapperFn.apply(x.float().contiguous(), self.bit_embeddings.weight)


def _encoder_param_grads(g_tot, msg, w1, b1, wc, base, res):
    g_w1, g_b1 = torch.zeros_like(w1), torch.zeros_like(b1)
    g_wc, g_bc = torch.zeros_like(wc), torch.zeros(4, dtype=torch.float32, device=w1.device)
    ops.secret_encoder_bwd(g_tot.float().contiguous(), msg, w1, b1, wc, g_w1, g_b1, g_wc, g_bc, base, res)
    return g_w1, g_b1, g_wc, g_bc


class _EncoderFn(torch.autograd.Function):
    """(x + c, c) with c = the resized secret map; backward = aq_secret_encoder_bwd (parameter gradients), dL/dx = dL/dx_out."""

    @staticmethod
    def forward(ctx, msg, w1, b1, wc, bc, x, base, res):
        xo, c = ops.secret_encoder_fwd(msg, w1, b1, wc, bc, x, (x.shape[2], x.shape[3]), base, res)
        ctx.save_for_backward(msg, w1, b1, wc)
        ctx.meta = (base, res)
        return xo, c

    @staticmethod
    def backward(ctx, g_xo, g_c):
        msg, w1, b1, wc = ctx.saved_tensors
        g_tot = g_c if g_xo is None else (g_xo if g_c is None else g_c + g_xo)
        grads = (None,) * 4 if g_tot is None else _encoder_param_grads(g_tot, msg, w1, b1, wc, *ctx.meta)
        return (None, *grads, g_xo, None, None)


class _EncoderMapFn(torch.autograd.Function):
    """`encode`: the secret map alone, at resolution x resolution."""

    @staticmethod
    def forward(ctx, msg, w1, b1, wc, bc, base, res):
        _, c = ops.secret_encoder_fwd(msg, w1, b1, wc, bc, None, (res, res), base, res)
        ctx.save_for_backward(msg, w1, b1, wc)
        ctx.meta = (base, res)
        return c

    @staticmethod
    def backward(ctx, g_c):
        msg, w1, b1, wc = ctx.saved_tensors
        return (None, *_encoder_param_grads(g_c, msg, w1, b1, wc, *ctx.meta), None, None)


class SecretEncoder(nn.Module):
    """Same parameter layout as the reference: secret_scaler.0 = Linear(secret_len, base_res^2), secret_scaler.5 =
    zero-initialised Conv2d(4, 4, 3, padding=1); indices 1-4 (SiLU, View, Repeat, Upsample) carry no parameters."""

    def __init__(self, secret_len, base_res=32, resolution=64) -> None:
        super().__init__()
        self.secret_len = secret_len
        self.base_res = base_res
        self.resolution = 2 ** int(np.log2(resolution))
        conv = nn.Conv2d(4, 4, 3, padding=1)
        for p in conv.parameters():
            p.detach().zero_()
        self.secret_scaler = nn.Sequential(nn.Linear(secret_len, base_res * base_res), nn.Identity(), nn.Identity(), nn.Identity(),
                                           nn.Identity(), conv)

    def _run(self, x, c):
        if not c.is_cuda:
            raise AqualoraError("SecretEncoder: CPU tensor passed; aqualora_b200 has no CPU fallback")
        lin, conv = self.secret_scaler[0], self.secret_scaler[5]
        for p in (lin.weight, lin.bias, conv.weight, conv.bias):
            if p.dtype != torch.float32:
                raise AqualoraError("SecretEncoder: parameters must be fp32 (the reference never casts this module)")
        msg = c.float().contiguous()
        if x is None:
            return None, _EncoderMapFn.apply(msg, lin.weight, lin.bias, conv.weight, conv.bias, self.base_res, self.resolution)
        return _EncoderFn.apply(msg, lin.weight, lin.bias, conv.weight, conv.bias, x.float(), self.base_res, self.resolution)

    def encode(self, x):
        return self._run(None, x)[1]

    def forward(self, x, c):
        xo, c_map = self._run(x, c)
        return xo.to(x.dtype), c_map
