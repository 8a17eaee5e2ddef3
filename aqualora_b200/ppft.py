"""Prior-preserving fine-tuning (PPFT) step around the fused watermark-LoRA kernels.

Mirrors the loop body of train/ppft_train.py:987-1068 on one GPU per process:

    msg -> scale = mapper(msg)                         (:989-990)      aq_mapper_fwd
    x_clean / x_wm = add_noise(z), add_noise(z + wm)   (:1010-1011)
    clean_pred = unet(x_clean, scale = 0).detach()     (:1026-1029)    LoRA skipped: scale 0 is bit-identical to the base op
    model_pred = unet(x_wm, scale)                     (:1032-1035)    192 fused projection+LoRA kernels
    loss = mse(model_pred, clean_pred); backward       (:1051-1058)    fused dX / weight-grad kernels
    DDP gradient allreduce over the LoRA (+mapper) gradients only      one NCCL allreduce of the flat fp32 buffer
    clip_grad_norm_(lora, 1.0); AdamW; zero_grad       (:1065-1068)    aq_flat_sumsq + aq_flat_clip_adamw

All trainable state lives in four flat fp32 buffers (param / grad / exp_avg / exp_avg_sq): the 384 LoRA matrices in
`unet_keys.json` order (down then up per target), then the mapper's bit embeddings.  The backward kernels accumulate
straight into the flat gradient buffer (`param._aq_grad` views), so the data-parallel exchange is ONE allreduce.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import checkpoint, lora_modules, ops
from .unet import UNet2DConditionModel, UNetConfig, lora_target_keys


@dataclass
class PPFTConfig:
    rank: int = 64
    msg_bits: int = 48
    learning_rate: float = 1e-4
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_weight_decay: float = 1e-2
    adam_epsilon: float = 1e-8
    max_grad_norm: float = 1.0
    lr_warmup_steps: int = 0
    max_train_steps: int = 1000
    lr_end: float = 0.0
    prediction_type: str = "epsilon"      # "v_prediction" for SD 2.x (train/ppft_train.py:1047-1049)
    scaling_factor: float = 0.18215       # vae.config.scaling_factor


def scaled_linear_alphas_cumprod(num_steps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    """DDPM `scaled_linear` schedule used by Stable Diffusion (scripts/lib/model_util.py:14-16)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_steps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(alphas_cumprod: torch.Tensor, x: torch.Tensor, noise: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """DDPMScheduler.add_noise: sqrt(a_t) x + sqrt(1 - a_t) noise."""
    a = alphas_cumprod.to(device=x.device, dtype=x.dtype)[t]
    return a.sqrt().view(-1, 1, 1, 1) * x + (1 - a).sqrt().view(-1, 1, 1, 1) * noise


def velocity_to_epsilon(alphas_cumprod: torch.Tensor, v: torch.Tensor, noisy: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """utils/cschedulers.py:56-72."""
    a = alphas_cumprod.to(device=t.device)[t]
    return (1 - a).sqrt()[:, None, None, None] * noisy + a.sqrt()[:, None, None, None] * v


def cosine_lr_factor(step: int, warmup: int, total: int, lr_end: float = 0.0, num_cycles: float = 0.5) -> float:
    """utils/misc.py:24-34 (get_cosine_schedule_with_warmup_lr_end)."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(lr_end, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


class FlatLoraState:
    """Flat fp32 buffers holding every trainable tensor; parameters are views, `_aq_grad` points into `grad`."""

    def __init__(self, lora_layers, mapper_emb: torch.Tensor, device):
        shapes = []
        for _, _, lora in lora_layers:
            shapes.append(lora.down.weight.shape)
            shapes.append(lora.up.weight.shape)
        self.n_lora = sum(math.prod(s) for s in shapes)
        self.n_lora_pad = (self.n_lora + 3) // 4 * 4
        self.n_mapper = mapper_emb.numel()
        total = self.n_lora_pad + (self.n_mapper + 3) // 4 * 4
        self.param = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.param)
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=device)
        off = 0
        self.params = []
        for _, _, lora in lora_layers:
            for lin in (lora.down, lora.up):
                n = lin.weight.numel()
                view = self.param[off:off + n].view(lin.weight.shape)
                view.copy_(lin.weight.detach().to(device=device, dtype=torch.float32))
                p = nn.Parameter(view, requires_grad=True)
                p._aq_grad = self.grad[off:off + n]
                lin.weight = p
                self.params.append(p)
                off += n
        self.mapper_off = self.n_lora_pad
        mview = self.param[self.mapper_off:self.mapper_off + self.n_mapper].view(mapper_emb.shape)
        mview.copy_(mapper_emb.detach().to(device=device, dtype=torch.float32))
        self.mapper_emb = mview
        self.mapper_grad = self.grad[self.mapper_off:self.mapper_off + self.n_mapper].view(mapper_emb.shape)

    @property
    def lora_param(self):
        return self.param[:self.n_lora_pad]

    def region(self, buf: torch.Tensor, which: str) -> torch.Tensor:
        return buf[:self.n_lora_pad] if which == "lora" else buf[self.mapper_off:]


class PPFTTrainer:
    """One process = one GPU.  `step()` runs one PPFT iteration on a synthetic or real latent batch."""

    def __init__(self, unet: UNet2DConditionModel, cfg: PPFTConfig, mapper_emb: torch.Tensor, device,
                 lora_up_std: Optional[float] = None, seed: int = 0):
        self.cfg = cfg
        self.device = torch.device(device)
        self.unet = unet
        unet.requires_grad_(False)
        self.keys = lora_target_keys(unet)
        self.lora_layers = lora_modules.inject_lora(unet, self.keys, cfg.rank)
        if lora_up_std is not None:
            # benchmarking / gradient-parity init: a zero `up` makes dDn and dscale vanish
            g = torch.Generator().manual_seed(seed)
            for _, _, lora in self.lora_layers:
                lora.up.weight.data.copy_(torch.randn(lora.up.weight.shape, generator=g) * lora_up_std)
        self.state = FlatLoraState(self.lora_layers, mapper_emb, self.device)
        self.alphas_cumprod = scaled_linear_alphas_cumprod().to(self.device)
        self.global_step = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.g_scale = None
        self._graph = None

    # -- pieces of the step ------------------------------------------------------------------
    def mapper(self, msg: torch.Tensor) -> torch.Tensor:
        """scale [B, r]: fp32 values rounded to bf16 (`.to(dtype=weight_dtype)`, train/ppft_train.py:990); a leaf whose
        gradient the projection kernels accumulate into `self.g_scale`."""
        scale = ops.mapper_fwd(msg, self.state.mapper_emb, round_bf16=True)
        if self.g_scale is None or self.g_scale.shape != scale.shape:
            self.g_scale = torch.zeros_like(scale)
        else:
            self.g_scale.zero_()
        scale.requires_grad_(True)
        scale._aq_grad = self.g_scale
        return scale

    def lr(self) -> float:
        c = self.cfg
        # accelerate steps the scheduler once per process (train/ppft_train.py:896-901 scales the horizon by num_processes)
        return c.learning_rate * cosine_lr_factor(self.global_step * self.world, c.lr_warmup_steps * self.world,
                                                  c.max_train_steps * self.world, c.lr_end)

    def forward_backward(self, latents, wm_latent, noise, timesteps, ctx, msg):
        c = self.cfg
        scale = self.mapper(msg)
        noisy = add_noise(self.alphas_cumprod, latents, noise, timesteps)
        noisy_wm = add_noise(self.alphas_cumprod, latents + wm_latent, noise, timesteps)
        with torch.no_grad(), lora_modules.lora_disabled():
            clean_pred = self.unet(noisy, timesteps, ctx).sample
        model_pred = self.unet(noisy_wm, timesteps, ctx, cross_attention_kwargs={"scale": scale}).sample
        if c.prediction_type == "v_prediction":
            model_pred = velocity_to_epsilon(self.alphas_cumprod, model_pred, noisy_wm, timesteps)
            clean_pred = velocity_to_epsilon(self.alphas_cumprod, clean_pred, noisy, timesteps)
        loss = F.mse_loss(model_pred.float(), clean_pred.float(), reduction="mean")
        loss.backward()
        lora_modules.flush_wgrad_queue()          # (the autograd engine's end-of-backward callback already did; explicit for clarity)
        ops.mapper_bwd(msg, self.g_scale, self.state.mapper_grad)
        return loss.detach()

    def exchange_gradients(self) -> float:
        """The only data-path collective of the PPFT step: ONE sum-allreduce of the flat LoRA (+mapper) gradient buffer over
        NVLink (what DDP does bucket by bucket for the 385 trainable tensors, train/ppft_train.py:906-912).  Returns the factor
        that turns the sum into DDP's mean; it is folded into the clip/AdamW kernel instead of a separate pass."""
        if self.world > 1:
            dist.all_reduce(self.state.grad)
        return 1.0 / self.world

    def optimizer_step(self):
        c, st = self.cfg, self.state
        gs = self.exchange_gradients()
        st.norm_sq.zero_()
        ops.flat_sumsq(st.region(st.grad, "lora"), st.norm_sq)       # clip_grad_norm_ covers the U-Net LoRA params only
        lr = self.lr()
        self.global_step += 1
        common = dict(grad_scale=gs, lr=lr, beta1=c.adam_beta1, beta2=c.adam_beta2, eps=c.adam_epsilon,
                      weight_decay=c.adam_weight_decay, step=self.global_step)
        ops.flat_clip_adamw(st.region(st.param, "lora"), st.region(st.grad, "lora"), st.region(st.exp_avg, "lora"),
                            st.region(st.exp_avg_sq, "lora"), st.norm_sq, max_norm=c.max_grad_norm, **common)
        ops.flat_clip_adamw(st.region(st.param, "mapper"), st.region(st.grad, "mapper"), st.region(st.exp_avg, "mapper"),
                            st.region(st.exp_avg_sq, "mapper"), st.norm_sq, max_norm=0.0, **common)
        lora_modules.invalidate_packed()                             # the bf16 operand copies are stale now ...
        lora_modules.refresh_packed(st.params)                       # ... and rebuilt by one launch over all 384 matrices

    def refresh_operands(self):
        """Rebuild the bf16 operand copies from the fp32 master parameters.  `optimizer_step` does it; call it after changing
        `state.param` any other way (checkpoint load) when stepping through the captured graph, which cannot notice the change."""
        lora_modules.invalidate_packed()
        lora_modules.refresh_packed(self.state.params)

    def step(self, latents, wm_latent, noise, timesteps, ctx, msg):
        loss = self.forward_backward(latents, wm_latent, noise, timesteps, ctx, msg)
        self.optimizer_step()
        return loss

    # -- checkpoint artefacts (train/ppft_train.py:699-748, :943-965, :1079-1103, :1203-1229) ---------------------------------
    def save(self, output_dir: str, msgdecoder: Optional[nn.Module] = None) -> None:
        """Final artefacts: pytorch_lora_weights.safetensors + mapper.pt (+ msgdecoder.pt) -- what scripts/create_wm_lora.py reads."""
        checkpoint.save_ppft_artifacts(output_dir, self.unet, self.keys, self.state.mapper_emb, msgdecoder)

    def load(self, input_dir: str, strict: bool = True) -> None:
        """Load LoRA weights (+ mapper.pt when present) written by `save` or by the reference's training script, in place into the flat
        parameter buffer, and rebuild the bf16 operand copies."""
        checkpoint.load_lora_into_unet(input_dir, self.unet, self.keys, strict=strict)
        mp = os.path.join(input_dir, "mapper.pt")
        if os.path.isdir(input_dir) and os.path.exists(mp):
            emb = torch.load(mp, map_location="cpu")["bit_embeddings.weight"]
            if tuple(emb.shape) != tuple(self.state.mapper_emb.shape):
                raise ValueError(f"mapper.pt holds {tuple(emb.shape)} embeddings, the trainer {tuple(self.state.mapper_emb.shape)}")
            with torch.no_grad():
                self.state.mapper_emb.copy_(emb.to(self.state.mapper_emb.device, torch.float32))
        if self.device.type == "cuda":
            self.refresh_operands()

    def save_state(self, output_dir: str, total_limit: Optional[int] = None, msgdecoder: Optional[nn.Module] = None) -> str:
        """`accelerator.save_state(<output_dir>/checkpoint-<global_step>)` with `--checkpoints_total_limit` rotation: the artefacts
        of `save` plus the optimizer moments and the step counter needed to resume."""
        os.makedirs(output_dir, exist_ok=True)
        checkpoint.rotate_checkpoints(output_dir, total_limit)
        path = os.path.join(output_dir, f"checkpoint-{self.global_step}")
        self.save(path, msgdecoder)
        st = self.state
        torch.save({"global_step": self.global_step, "n_lora": st.n_lora, "n_mapper": st.n_mapper, "rank": self.cfg.rank,
                    "exp_avg": st.exp_avg.detach().cpu().clone(), "exp_avg_sq": st.exp_avg_sq.detach().cpu().clone()},
                   os.path.join(path, "trainer_state.pt"))
        return path

    def load_state(self, path: str) -> int:
        """Resume from a `save_state` directory (or `latest` inside an output dir: train/ppft_train.py:943-965).  Returns the step."""
        if os.path.basename(os.path.normpath(path)) == "latest":
            found = checkpoint.latest_checkpoint(os.path.dirname(os.path.normpath(path)))
            if found is None:
                raise FileNotFoundError(f"no checkpoint-* directory under {os.path.dirname(os.path.normpath(path))}")
            path = found
        self.load(path)
        ts = torch.load(os.path.join(path, "trainer_state.pt"), map_location="cpu")
        st = self.state
        if ts["n_lora"] != st.n_lora or ts["n_mapper"] != st.n_mapper:
            raise ValueError("trainer_state.pt belongs to a different LoRA configuration")
        st.exp_avg.copy_(ts["exp_avg"].to(st.exp_avg.device))
        st.exp_avg_sq.copy_(ts["exp_avg_sq"].to(st.exp_avg_sq.device))
        self.global_step = int(ts["global_step"])
        return self.global_step

    # -- CUDA-graph replay of the forward + backward ------------------------------------------------------------------
    def capture(self, fn: Callable, example_inputs, warmup_steps: int = 2) -> int:
        """Record `fn(*inputs) -> loss` -- everything of a step up to the gradient exchange, ending in `forward_backward` -- into
        a CUDA graph over static copies of `example_inputs` (fixed shapes and dtypes).  A PPFT step issues ~10 k kernels; replayed
        as one graph the host stops being the bottleneck once the glue kernels have shortened the device time.  The gradient
        all-reduce and the clip / AdamW kernels stay outside (their learning rate and bias-correction step are launch arguments).
        Returns the number of this library's kernel launches inside the graph (bench.py counts them per replay)."""
        from . import _lib

        stream = torch.cuda.Stream(device=self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        self._static_in = [x.clone() for x in example_inputs]
        with torch.cuda.stream(stream):
            for _ in range(max(1, warmup_steps)):      # handles, cuDNN plans and workspaces of the capture stream:
                fn(*self._static_in)                   # forward + backward only -- the warm-up must not train on the example
        torch.cuda.current_stream(self.device).wait_stream(stream)
        torch.cuda.synchronize(self.device)
        self.state.grad.zero_()                        # batch (no optimizer step, no moment / step-counter / schedule change)
        if self.g_scale is not None:
            self.g_scale.zero_()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.load().aq_launch_count()
        with torch.cuda.graph(graph, stream=stream):
            self._static_loss = fn(*self._static_in)
        self._graph_launches = int(_lib.load().aq_launch_count() - n0)
        self._graph = graph
        self.state.grad.zero_()                        # a capture records kernels without running them, but stay defensive:
        return self._graph_launches                    # the first replay must start from zero gradients

    @property
    def static_inputs(self):
        """The captured graph's input buffers (same order as the `example_inputs` of `capture`): write the next batch here
        (e.g. `copy_` straight from pinned host memory) and call `step_graphed(copy_inputs=False)`."""
        if self._graph is None:
            raise RuntimeError("PPFTTrainer.capture() has not been called")
        return self._static_in

    def step_graphed(self, *inputs, copy_inputs: bool = True):
        """One PPFT step through the captured graph: inputs -> static buffers, replay, gradient exchange + clip + AdamW."""
        if self._graph is None:
            raise RuntimeError("PPFTTrainer.capture() has not been called")
        if copy_inputs:
            for dst, src in zip(self._static_in, inputs):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.optimizer_step()
        return self._static_loss


def build_unet(cfg: UNetConfig, device, dtype=torch.bfloat16, seed: int = 0) -> UNet2DConditionModel:
    """Random-init U-Net of the named architecture (no checkpoints exist offline), cast like ppft_train.py:569-581."""
    torch.manual_seed(seed)
    with torch.device(device):
        unet = UNet2DConditionModel(cfg)
    unet = unet.to(dtype=dtype)
    if torch.device(device).type == "cuda":
        unet = unet.to(memory_format=torch.channels_last)
    unet.requires_grad_(False)
    unet.eval()
    return unet
