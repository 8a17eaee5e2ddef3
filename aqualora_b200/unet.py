"""Stable-Diffusion U-Net harness around the watermark-LoRA projections (host plumbing, plain PyTorch).

The hot path of this repository is the 192 LoRA-target projections (aqualora_b200/lora_modules.py -> CUDA).  This file
is the CALLER on either side of it: a compact `UNet2DConditionModel` whose module paths and state-dict keys are the
diffusers ones, so that `utils/unet_keys.json` (train/ppft_train.py:620-689) resolves on it and checkpoints /
`pytorch_lora_weights.safetensors` keys (train/ppft_train.py:443-471) line up.  Everything that is not a LoRA target
(3x3 convolutions, attention core, residual adds) is a library call; GroupNorm(+SiLU), LayerNorm and GEGLU run as CUDA
kernels of this repository on the GPU path (aqualora_b200/unet_ops.py).

Topology follows the SD 1.5 / 2.1 configs quoted in scripts/lib/original_unet.py:22-106; the golden test loads the
same procedurally generated weights into the reference's vendored U-Net and into this one and compares outputs.
Activations are kept channels_last so a Transformer2D block sees [B, H*W, C] token-major memory without copies.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .lora_modules import (AquaLoRAAttnProcessor, LoRACompatibleConv, LoRACompatibleLinear, conv1x1_with_residual, linear_with_residual,
                           precompute_cross_kv)
from .unet_ops import geglu, group_norm_nhwc, layer_norm, residual_add_bias


import os

LIBRARY_GLUE = False     # bench.py's PyTorch-eager GPU baseline sets this: every norm / activation stays a library op
# Fold the three residual adds of a transformer block (and the one after proj_out) into the projection's tile epilogue
# (aq_lora_linear_fwd_residual).  Measured on B200 (profiles/r02_bench_v3_residual_fused.json): the step gains 1.0 ms (77.3 -> 76.3 ms,
# 96 elementwise adds fewer per step), but the epilogue's row-strided reads of a residual that sits in HBM cost the fused forward
# GEMMs 0.9 ms and the plain ones 0.65 ms per step (roofline fraction 0.62 -> 0.54).  Off by default until the residual tile is
# TMA-prefetched into the staging buffer ahead of the accumulator wait; AQ_FUSE_RESIDUAL=1 turns it on.
FUSE_RESIDUAL = os.environ.get("AQ_FUSE_RESIDUAL", "0") == "1"


def _fused_glue(x: torch.Tensor, *params: torch.Tensor) -> bool:
    """The glue kernels (aq_group_norm_nhwc_*, aq_geglu_*) serve the frozen bf16 U-Net on the GPU; the fp32 CPU copy of this
    module tree that tests / the CPU baseline patch with the oracle keeps the library ops."""
    return not LIBRARY_GLUE and x.is_cuda and x.dtype == torch.bfloat16 and not any(p.requires_grad for p in params)


def _norm_act(norm: nn.GroupNorm, x: torch.Tensor, silu: bool, add_bc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[silu](norm(x + add_bc[:, :, None, None])) -- scripts/lib/original_unet.py:440-453, :826, :1416."""
    if _fused_glue(x, norm.weight, norm.bias) and (add_bc is None or not add_bc.requires_grad):
        return group_norm_nhwc(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu, add_bc)
    if add_bc is not None:
        x = x + add_bc[:, :, None, None]
    y = norm(x)
    return F.silu(y) if silu else y


@dataclass
class UNetConfig:
    sample_size: int = 64
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Sequence[int] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: Sequence[int] | int = 8      # SD1.5: number of heads (diffusers' historical misnomer)
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    upcast_attention: bool = False
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    down_has_attn: Sequence[bool] = (True, True, True, False)
    time_embed_dim: int = field(default=0)

    def heads(self, level: int) -> int:
        if isinstance(self.attention_head_dim, int):
            return self.attention_head_dim
        return self.attention_head_dim[level]

    @staticmethod
    def sd15(sample_size: int = 64) -> "UNetConfig":
        return UNetConfig(sample_size=sample_size)

    @staticmethod
    def sd21(sample_size: int = 96) -> "UNetConfig":
        return UNetConfig(sample_size=sample_size, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                          use_linear_projection=True, upcast_attention=True)

    @staticmethod
    def tiny(sample_size: int = 16) -> "UNetConfig":
        """Same topology, narrow channels: CPU-sized for tests (not a BASELINE config)."""
        return UNetConfig(sample_size=sample_size, block_out_channels=(32, 64, 128, 128), attention_head_dim=4,
                          cross_attention_dim=64, norm_num_groups=8)


def timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """Sinusoidal embedding, flip_sin_to_cos=True, freq_shift=0 (scripts/lib/original_unet.py:323-362)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


def _token_norm(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """BasicTransformerBlock.norm1/2/3 (scripts/lib/original_unet.py:732-806)."""
    if _fused_glue(x, norm.weight, norm.bias):
        return layer_norm(x, norm.weight, norm.bias, norm.eps)
    return norm(x)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        frozen = [self.conv1.weight, self.conv1.bias, self.conv2.bias, self.norm1.weight, self.norm2.weight]
        if self.conv_shortcut is not None:
            frozen.append(self.conv_shortcut.bias)
        if _fused_glue(x, *frozen):
            # no broadcast-bias pass after the convolutions: conv1's bias rides on the time embedding that norm2's kernel adds,
            # conv2's (and the shortcut's) on the closing residual add
            h = F.conv2d(_norm_act(self.norm1, x, True), self.conv1.weight, None, padding=1)
            t = self.time_emb_proj(F.silu(temb)) + self.conv1.bias
            h = F.conv2d(_norm_act(self.norm2, h, True, add_bc=t), self.conv2.weight, None, padding=1)
            bias = self.conv2.bias
            if self.conv_shortcut is not None:
                x = F.conv2d(x, self.conv_shortcut.weight, None)
                bias = self._folded_bias()
            return residual_add_bias(x, h, bias)
        h = self.conv1(_norm_act(self.norm1, x, True))
        h = self.conv2(_norm_act(self.norm2, h, True, add_bc=self.time_emb_proj(F.silu(temb))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h

    def _folded_bias(self):
        """conv2.bias + conv_shortcut.bias, cached (both frozen)."""
        key = (self.conv2.bias._version, self.conv_shortcut.bias._version, self.conv2.bias.data_ptr())
        hit = getattr(self, "_aq_bias_sum", None)
        if hit is None or hit[0] != key:
            hit = (key, (self.conv2.bias.float() + self.conv_shortcut.bias.float()).to(self.conv2.bias.dtype))
            self._aq_bias_sum = hit
        return hit[1]


class Attention(nn.Module):
    """diffusers `Attention` / kohya `CrossAttention`: q/k/v/out are the LoRA targets (utils/unet_keys.json)."""

    def __init__(self, query_dim: int, context_dim: Optional[int], heads: int, upcast: bool):
        super().__init__()
        context_dim = context_dim or query_dim
        self.heads = heads
        self.upcast = upcast
        self.to_q = LoRACompatibleLinear(query_dim, query_dim, bias=False)
        self.to_k = LoRACompatibleLinear(context_dim, query_dim, bias=False)
        self.to_v = LoRACompatibleLinear(context_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([LoRACompatibleLinear(query_dim, query_dim)])
        self.processor = AquaLoRAAttnProcessor()

    def forward(self, x, context=None, scale=1.0, residual=None):
        """`residual`: the stream the caller adds to the attention output (x + attn(norm(x))): folded into to_out's epilogue."""
        if residual is None:
            return self.processor(self, x, encoder_hidden_states=context, scale=scale)
        return self.processor(self, x, encoder_hidden_states=context, scale=scale, residual=residual)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = LoRACompatibleLinear(dim_in, dim_out * 2)

    def forward(self, x, scale=1.0):
        p = self.proj(x, scale)
        if _fused_glue(p):
            return geglu(p)
        h, gate = p.chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), LoRACompatibleLinear(dim * 4, dim)])

    def forward(self, x, scale=1.0, residual=None):
        h = self.net[0](x, scale)
        if residual is None:
            return self.net[2](h, scale)
        return linear_with_residual(self.net[2], h, scale, residual)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, context_dim: int, upcast: bool):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, upcast)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, context_dim, heads, upcast)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context, scale=1.0):
        if FUSE_RESIDUAL and x.is_cuda:
            # the three residual adds ride in the epilogue of to_out / to_out / ff.net.2 (aq_lora_linear_fwd_residual)
            x = self.attn1(_token_norm(self.norm1, x), None, scale, residual=x)
            x = self.attn2(_token_norm(self.norm2, x), context, scale, residual=x)
            return self.ff(_token_norm(self.norm3, x), scale, residual=x)
        x = self.attn1(_token_norm(self.norm1, x), None, scale) + x
        x = self.attn2(_token_norm(self.norm2, x), context, scale) + x
        return self.ff(_token_norm(self.norm3, x), scale) + x


class Transformer2DModel(nn.Module):
    def __init__(self, channels: int, heads: int, context_dim: int, groups: int, linear_proj: bool, upcast: bool):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        if linear_proj:
            self.proj_in = LoRACompatibleLinear(channels, channels)
            self.proj_out = LoRACompatibleLinear(channels, channels)
        else:
            self.proj_in = LoRACompatibleConv(channels, channels, kernel_size=1)
            self.proj_out = LoRACompatibleConv(channels, channels, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, context_dim, upcast)])

    def forward(self, x, context, scale=1.0):
        B, C, H, W = x.shape
        res = x
        h = _norm_act(self.norm, x, False)
        if self.linear_proj:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C), scale)
        else:
            h = self.proj_in(h, scale).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, context, scale)
        if not (FUSE_RESIDUAL and x.is_cuda):
            if self.linear_proj:
                h = self.proj_out(h, scale).reshape(B, H, W, C).permute(0, 3, 1, 2)
            else:
                h = self.proj_out(h.reshape(B, H, W, C).permute(0, 3, 1, 2), scale)
            return h + res
        if self.linear_proj:
            if res.is_cuda and res.is_contiguous(memory_format=torch.channels_last):
                return linear_with_residual(self.proj_out, h, scale, res.permute(0, 2, 3, 1).reshape(B, H * W, C)).reshape(B, H, W, C).permute(0, 3, 1, 2)
            h = self.proj_out(h, scale).reshape(B, H, W, C).permute(0, 3, 1, 2)
            return h + res
        return conv1x1_with_residual(self.proj_out, h.reshape(B, H, W, C).permute(0, 3, 1, 2), scale, res)


class Downsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x, size=None):
        if size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=size, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin: int, cout: int, level: int, has_attn: bool, add_down: bool):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for j in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(cout, cfg.heads(level), cfg.cross_attention_dim, cfg.norm_num_groups,
                                                          cfg.use_linear_projection, cfg.upcast_attention))
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, context, scale):
        outs = []
        for j, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, context, scale)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, ch: int):
        super().__init__()
        mk = lambda: ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
        self.resnets = nn.ModuleList([mk(), mk()])
        # the vendored reference does not forward upcast_attention to the mid block (original_unet.py:1372-1377)
        self.attentions = nn.ModuleList([Transformer2DModel(ch, cfg.heads(len(cfg.block_out_channels) - 1), cfg.cross_attention_dim,
                                                            cfg.norm_num_groups, cfg.use_linear_projection, False)])

    def forward(self, x, temb, context, scale):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context, scale)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin: int, cout: int, cprev: int, level: int, has_attn: bool, add_up: bool):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        n = cfg.layers_per_block + 1
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = cprev if j == 0 else cout
            self.resnets.append(ResnetBlock2D(rin + skip, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(cout, cfg.heads(level), cfg.cross_attention_dim, cfg.norm_num_groups,
                                                          cfg.use_linear_projection, cfg.upcast_attention))
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, context, scale, up_size=None):
        for j, res in enumerate(self.resnets):
            x = res(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[j](x, context, scale)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x, up_size)
        return x


class UNetOutput:
    def __init__(self, sample):
        self.sample = sample


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        ch = list(cfg.block_out_channels)
        cfg.time_embed_dim = ch[0] * 4
        self.cfg = cfg
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], cfg.time_embed_dim)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i in range(len(ch)):
            cin, cout = cout, ch[i]
            self.down_blocks.append(DownBlock(cfg, cin, cout, i, cfg.down_has_attn[i], add_down=i < len(ch) - 1))
        self.mid_block = MidBlock(cfg, ch[-1])
        self.up_blocks = nn.ModuleList()
        rev = ch[::-1]
        rev_attn = list(cfg.down_has_attn)[::-1]
        cout = rev[0]
        for i in range(len(ch)):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(UpBlock(cfg, cin, cout, cprev, len(ch) - 1 - i, rev_attn[i], add_up=i < len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)
        self._cross_attentions = None

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, cross_attention_kwargs=None,
                return_dict: bool = True):
        scale = 1.0 if cross_attention_kwargs is None else cross_attention_kwargs.get("scale", 1.0)
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=sample.device)
        timestep = timestep.reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(timestep, self.cfg.block_out_channels[0]).to(self.dtype))
        if sample.is_cuda:
            sample = sample.contiguous(memory_format=torch.channels_last)
        if sample.is_cuda:
            # the 2 x 16 cross-attention K / V projections all read the text context: one grouped launch up front
            if self._cross_attentions is None:
                self._cross_attentions = [m.attn2 for m in self.modules() if isinstance(m, BasicTransformerBlock)]
            precompute_cross_kv(self._cross_attentions, encoder_hidden_states.to(self.dtype), scale)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states, scale)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states, scale)
        n_up = len(self.up_blocks)
        odd = any(s % (2 ** (n_up - 1)) for s in sample.shape[-2:])
        for i, blk in enumerate(self.up_blocks):
            n_res = len(blk.resnets)
            up_size = None
            if odd and i < n_up - 1:
                up_size = skips[-n_res - 1].shape[2:]
            x = blk(x, skips, temb, encoder_hidden_states, scale, up_size)
        x = self.conv_out(_norm_act(self.conv_norm_out, x, True))
        return UNetOutput(x) if return_dict else (x,)


def lora_target_keys(unet: UNet2DConditionModel) -> list[str]:
    """The ordered list the reference ships as utils/unet_keys.json (192 entries for SD 1.5 / 2.1): every
    Transformer2D proj_in/proj_out, attn{1,2}.to_{k,out.0,q,v} and ff.net.{0.proj,2}, block by block in the file's order
    (alphabetical inside a transformer)."""
    keys: list[str] = []

    def add(prefix: str, blocks):
        for i, blk in enumerate(blocks):
            if getattr(blk, "attentions", None) is None:
                continue
            for j in range(len(blk.attentions)):
                base = f"{prefix}.{i}.attentions.{j}"
                keys.extend([f"{base}.proj_in", f"{base}.proj_out"])
                tb = f"{base}.transformer_blocks.0"
                for attn in ("attn1", "attn2"):
                    keys.extend([f"{tb}.{attn}.to_k", f"{tb}.{attn}.to_out.0", f"{tb}.{attn}.to_q", f"{tb}.{attn}.to_v"])
                keys.extend([f"{tb}.ff.net.0.proj", f"{tb}.ff.net.2"])

    add("down_blocks", unet.down_blocks)
    base = "mid_block.attentions.0"
    mid: list[str] = [f"{base}.proj_in", f"{base}.proj_out"]
    tb = f"{base}.transformer_blocks.0"
    for attn in ("attn1", "attn2"):
        mid.extend([f"{tb}.{attn}.to_k", f"{tb}.{attn}.to_out.0", f"{tb}.{attn}.to_q", f"{tb}.{attn}.to_v"])
    mid.extend([f"{tb}.ff.net.0.proj", f"{tb}.ff.net.2"])
    keys.extend(mid)
    add("up_blocks", unet.up_blocks)
    return keys
