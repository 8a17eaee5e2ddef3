"""Monkey-patch forwards that run the oracle restatement on the in-repo containers, mirroring how the reference
patches diffusers' modules (train/ppft_train.py:681-689).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`patch_with_oracle(unet)` turns aqualora_b200.unet.UNet2DConditionModel (host plumbing, plain PyTorch) into the CPU
reference path: PyTorch-eager U-Net + the reference's unfused LoRA op sequence.
"""
from __future__ import annotations

import types

import torch.nn as nn
import torch.nn.functional as F

from . import lora_oracle as O


def oracle_lora_linear_layer_forward(self, hidden_states, scale=1.0):
    return O.lora_linear_layer_forward(hidden_states, self.down.weight, self.up.weight, scale, self.network_alpha, self.rank)


def oracle_lora_conv2d_layer_forward(self, hidden_states, scale=1.0):
    return O.lora_conv2d_layer_forward(hidden_states, self.down.weight, self.up.weight, scale, self.network_alpha, self.rank,
                                       self.down.stride, self.down.padding)


def oracle_compatible_linear_forward(self, hidden_states, scale=1.0):
    out = F.linear(hidden_states, self.weight, self.bias)
    if self.lora_layer is None:
        return out
    return out + self.lora_layer(hidden_states, scale)


def oracle_compatible_conv_forward(self, hidden_states, scale=1.0):
    out = F.conv2d(hidden_states, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
    if self.lora_layer is None:
        return out
    return out + self.lora_layer(hidden_states, scale)


def patch_with_oracle(root: nn.Module) -> int:
    n = 0
    for _, m in root.named_modules():
        if isinstance(m, nn.Conv2d) and hasattr(m, "lora_layer"):
            m.forward = types.MethodType(oracle_compatible_conv_forward, m)
            if m.lora_layer is not None:
                m.lora_layer.forward = types.MethodType(oracle_lora_conv2d_layer_forward, m.lora_layer)
            n += 1
        elif isinstance(m, nn.Linear) and hasattr(m, "lora_layer"):
            m.forward = types.MethodType(oracle_compatible_linear_forward, m)
            if m.lora_layer is not None:
                m.lora_layer.forward = types.MethodType(oracle_lora_linear_layer_forward, m.lora_layer)
            n += 1
    return n
