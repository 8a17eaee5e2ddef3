"""CPU oracle for hot path (i): the watermark-LoRA projections.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, autograd) restatement of the reference's arithmetic, one function per reference function,
each citing the lines it follows (paths relative to the AquaLoRA repository).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product (aqualora_b200/) never does.

Parity status: PINNED.  tests/golden/lora_*.pt were produced by importing the reference's own
utils/lora_modules.py unchanged (tools/gen_golden.py, stubbing only the absent `diffusers` names) and
tests/test_oracle_golden.py checks this restatement against them bit-for-bit in fp32.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def lora_linear_layer_forward(hidden_states, down_weight, up_weight, scale=1.0, network_alpha=None, rank=None):
    """utils/lora_modules.py:9-26 (CustomLoRALinearLayerforward).

    down = Linear(din, r, bias=False), up = Linear(r, dout, bias=False).  A tensor `scale` [B, r] is applied as a
    per-sample diagonal between down and up (:15-17); a float `scale` multiplies the result (:24-25).
    """
    orig_dtype = hidden_states.dtype
    dtype = down_weight.dtype
    down_hidden_states = F.linear(hidden_states.to(dtype), down_weight)          # :13
    if isinstance(scale, torch.Tensor):
        mid = torch.diag_embed(scale)                                            # :16
        down_hidden_states = down_hidden_states @ mid                            # :17
    up_hidden_states = F.linear(down_hidden_states, up_weight)                   # :19
    if network_alpha is not None:
        up_hidden_states = up_hidden_states * (network_alpha / rank)             # :21-22
    if isinstance(scale, float):
        return scale * up_hidden_states.to(orig_dtype)                           # :24-25
    return up_hidden_states.to(orig_dtype)                                       # :26


def lora_conv2d_layer_forward(hidden_states, down_weight, up_weight, scale=1.0, network_alpha=None, rank=None,
                              stride=1, padding=0):
    """utils/lora_modules.py:28-44 (CustomLoRAConv2dLayerforward); down is a kxk conv, up a 1x1 conv, no biases."""
    orig_dtype = hidden_states.dtype
    dtype = down_weight.dtype
    down_hidden_states = F.conv2d(hidden_states.to(dtype), down_weight, None, stride, padding)   # :32
    if isinstance(scale, torch.Tensor):
        down_hidden_states = down_hidden_states * scale[:, :, None, None]                        # :35
    up_hidden_states = F.conv2d(down_hidden_states, up_weight)                                   # :37
    if network_alpha is not None:
        up_hidden_states = up_hidden_states * (network_alpha / rank)                             # :39-40
    if isinstance(scale, float):
        return scale * up_hidden_states.to(orig_dtype)
    return up_hidden_states.to(orig_dtype)


def lora_compatible_linear_forward(hidden_states, weight, bias, lora=None, scale=1.0):
    """utils/lora_modules.py:56-62.  `lora` is None or a dict(down=, up=, network_alpha=, rank=)."""
    out = F.linear(hidden_states, weight, bias)
    if lora is None:
        return out                                                                               # :57-59
    return out + lora_linear_layer_forward(hidden_states, lora["down"], lora["up"], scale,
                                           lora.get("network_alpha"), lora.get("rank"))          # :61


def lora_compatible_conv_forward(hidden_states, weight, bias, lora=None, scale=1.0, stride=1, padding=0):
    """utils/lora_modules.py:46-54."""
    out = F.conv2d(hidden_states, weight, bias, stride, padding)
    if lora is None:
        return out                                                                               # :47-52
    return out + lora_conv2d_layer_forward(hidden_states, lora["down"], lora["up"], scale, lora.get("network_alpha"),
                                           lora.get("rank"), stride, padding)                    # :54


def mapper_forward(msg, bit_embeddings):
    """utils/models.py:110-115 (MapperNet.forward): 1 + sum_i msg[b, i] * E[i, :] / sqrt(bits)."""
    bits = bit_embeddings.shape[0]
    encoded = bit_embeddings[None, :, :] * msg[:, :, None]
    return encoded.sum(dim=1) / torch.sqrt(torch.tensor(bits).float()) + 1.0


def mapper_init(bits, rank, std=1.0, generator=None):
    """utils/models.py:102-108: orthogonal init, rows normalised to unit std, times std."""
    w = torch.empty(bits, rank)
    torch.nn.init.orthogonal_(w, generator=generator)
    w = w / w.std(dim=1, keepdim=True)
    return w * std


def closed_form_linear(x, weight, bias, down, up, scale, alpha_over_rank=1.0):
    """The closed form stated in SURVEY.md 8(a) and DESIGN.md, used to cross-check the restatement above:
    Y = X W^T + b + a * ((X D^T) (.) s[:, None, :]) U^T for x [B, N, din], scale [B, r]."""
    h = x @ down.t()
    hs = h * scale[:, None, :]
    y = x @ weight.t()
    if bias is not None:
        y = y + bias
    return y + alpha_over_rank * (hs @ up.t())


def closed_form_linear_grads(x, weight, down, up, scale, gy, alpha_over_rank=1.0):
    """dX, dD, dU, ds of closed_form_linear (fp32 algebra, no autograd)."""
    a = alpha_over_rank
    h = x @ down.t()                      # [B, N, r]
    hs = h * scale[:, None, :]
    dhs = a * (gy @ up)                   # [B, N, r]
    d_up = a * torch.einsum("bno,bnr->or", gy, hs)
    d_scale = (dhs * h).sum(dim=1)        # [B, r]
    dh = dhs * scale[:, None, :]
    d_down = torch.einsum("bnr,bni->ri", dh, x)
    dx = gy @ weight + dh @ down
    return dx, d_down, d_up, d_scale


def bf16_autocast_linear(x_bf16, weight_bf16, bias_bf16, down_f32, up_f32, scale_bf16):
    """What the reference executes under `accelerate --mixed_precision bf16` (train/ppft_train.py:569-581, :990,
    :1026-1035): every GEMM takes bf16 operands with fp32 accumulation and rounds its output to bf16; the LoRA
    master weights are fp32 and are cast per call by autocast.  Emulated on CPU in fp32 with explicit rounding."""
    def r16(t):
        return t.to(torch.bfloat16).float()

    x = x_bf16.float()
    base = r16(x @ weight_bf16.float().t() + (0 if bias_bf16 is None else bias_bf16.float()))
    h = r16(x @ r16(down_f32).t())
    hs = r16(h * scale_bf16.float()[:, None, :])
    u = r16(hs @ r16(up_f32).t())
    return r16(base + u), h


def sqrt_bits(bits: int) -> float:
    return math.sqrt(float(bits))
