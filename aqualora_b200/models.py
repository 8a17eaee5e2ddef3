"""Drop-in counterparts of the reference's `utils/models.py` classes (same constructor signatures, attribute names and
state-dict keys), backed by the sm_100a kernels.  CPU tensors raise: there is no PyTorch fallback.

    MapperNet(input_size, output_size, std=1.)          utils/models.py:98-115
    SecretEncoder(secret_len, base_res=32, resolution=64)  utils/models.py:51-81   (inference; the PPFT loop calls it
                                                           under no_grad, train/ppft_train.py:994-996)
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
        hw = (self.resolution, self.resolution) if x is None else (x.shape[2], x.shape[3])
        with torch.no_grad():
            return ops.secret_encoder_fwd(c.float(), lin.weight.float(), lin.bias.float(), conv.weight.float(), conv.bias.float(),
                                          None if x is None else x.float(), hw, self.base_res, self.resolution)

    def encode(self, x):
        return self._run(None, x)[1]

    def forward(self, x, c):
        xo, c_map = self._run(x, c)
        return xo.to(x.dtype), c_map
