"""Stage-1 latent-watermark pretraining step (train/latent_wm_pretrain.py:164-217) on the GPU.

    latents = vae.encode(img).detach()                               :171    (third party, frozen: passed in)
    _, wm = sec_encoder(latents, msg)                                :174    aq_secret_encoder_fwd / _bwd
    z_w = gen_combined_latents(latents, wm, scale)                   :133-149, :176-179
    clean = vae.decode(latents).detach(); wmimg = vae.decode(z_w)    :180-181 (third party, frozen: passed in)
    lpips(clean, wmimg), PRVL_loss(clean, wmimg)                     :182-183 (LPIPS third party; PRVL: csrc/losses.cu)
    wmimg = noiser([wmimg, None], p)[0]                              :185-188 csrc/noise.cu forward + input-gradient kernels
    logits = sec_decoder(wmimg)                                      :190     SecretDecoder.train() (batch-stat BN, library convolutions)
    msgloss = BCE-with-logits(logits, one_hot(msg))                  :193-195 csrc/losses.cu
    loss schedule :206-214 ; backward :216 ; optimizer.step :217     (the caller owns the optimizer)

The VAE and LPIPS networks are frozen third-party models with no weights offline: they come in as callables.
"""
from __future__ import annotations

import random
from typing import Callable, Optional, Sequence

import torch
import torch.nn.functional as F

from . import losses


def draw_cornerfy(rng: random.Random):
    """The draws of gen_combined_latents in the reference's order (train/latent_wm_pretrain.py:134-137)."""
    cornerfy = rng.choice([True, False, False, False])
    hs, ws = (rng.uniform(1.0, 2.0), rng.uniform(1.0, 2.0)) if cornerfy else (1.0, 1.0)
    return cornerfy, hs, ws


def gen_combined_latents(latents: torch.Tensor, wm_latent: torch.Tensor, scale: float = 1.0, cornerfy: bool = False,
                         height_scale: float = 1.0, width_scale: float = 1.0) -> torch.Tensor:
    """train/latent_wm_pretrain.py:133-149 with the random choices passed in (a [B, 4, 64, 64] tensor: host-level glue, not a kernel):
    with `cornerfy`, the four corner quadrants of the watermark residual are pasted into the corners of a zero canvas enlarged by
    (height_scale, width_scale), which is resized back; latents + residual * scale."""
    if cornerfy:
        h, w = wm_latent.shape[2], wm_latent.shape[3]
        t = F.interpolate(torch.zeros_like(latents), scale_factor=(height_scale, width_scale), mode="bilinear")
        t[:, :, :h // 2, :w // 2] = wm_latent[:, :, :h // 2, :w // 2]
        t[:, :, :h // 2, -w // 2:] = wm_latent[:, :, :h // 2, -w // 2:]
        t[:, :, -h // 2:, :w // 2] = wm_latent[:, :, -h // 2:, :w // 2]
        t[:, :, -h // 2:, -w // 2:] = wm_latent[:, :, -h // 2:, -w // 2:]
        wm = F.interpolate(t, size=(h, w), mode="bilinear")
    else:
        wm = wm_latent
    return latents + wm * scale


def pretrain_step(sec_encoder, sec_decoder, vae_encode: Callable, vae_decode: Callable, lpips_fn: Callable, noiser, image: torch.Tensor,
                  msg: torch.Tensor, rng: random.Random, noise_probs: Sequence[float], warmup: bool, stage: int,
                  layer_override: Optional[Callable] = None):
    """Loop body of train/latent_wm_pretrain.py:171-216 up to and including `loss.backward()`.
    `stage`: 0 message loss only, 1 lpips + message, 2 5 lpips + message + 1.5 PRVL (:206-214); `warmup` forces the message loss and
    the 0.03 residual scale (:176-177, :206).  `layer_override(img) -> img` replaces the Noiser draw (tests pin the layer)."""
    latents = vae_encode(image).detach()
    _, wm_latent = sec_encoder(latents, msg.float())
    cornerfy, hs, ws = draw_cornerfy(rng)
    wl = gen_combined_latents(latents, wm_latent, 0.03 if warmup else 1.0, cornerfy, hs, ws)
    clean_image = vae_decode(latents).detach()
    wm_image = vae_decode(wl)
    lp = lpips_fn(clean_image, wm_image)
    prvl = losses.PRVL_loss(clean_image, wm_image)
    distorted = layer_override(wm_image) if layer_override is not None else noiser([wm_image, None], list(noise_probs))[0]
    reveal = sec_decoder(distorted)
    labels = F.one_hot(msg.long(), num_classes=2).float()
    msgloss = losses.binary_cross_entropy_with_logits(reveal, labels)
    if warmup or stage == 0:
        loss = msgloss
    elif stage == 1:
        loss = lp + msgloss
    else:
        loss = lp * 5 + msgloss * 1.0 + prvl * 1.5
    loss.backward()
    return {"loss": loss.detach(), "msgloss": msgloss.detach(), "lpips": lp.detach(), "prvl": prvl.detach(), "reveal": reveal.detach(),
            "wm_image": wm_image.detach(), "cornerfy": cornerfy}


def checkpoint_dict(sec_encoder, sec_decoder) -> dict:
    """train/latent_wm_pretrain.py:246-249 (read back at train/ppft_train.py:550-554)."""
    return {"sec_decoder": sec_decoder.state_dict(), "sec_encoder": sec_encoder.state_dict()}


def decoder_finetune_step(msgdecoder, images01: torch.Tensor, msg: torch.Tensor, distort: Optional[Callable] = None):
    """Stage-3 robustness fine-tuning of the decoder, loop body of train/rob_enhance_finetune.py:1020-1038 after the (third-party)
    sampling pipeline produced `images01` [B, 3, H, W] in [0, 1]: distort (utils/noise_layers/noiser.py:46-71 `distorsion_unit`
    mixture, passed in as a callable on [0, 1] images), map to [-1, 1], detach, decode in train mode, BCE against one_hot(msg),
    backward.  Returns (loss, validation accuracy) as the reference logs them."""
    x = images01.float()
    if distort is not None:
        x = distort(x)
    x = (x * 2 - 1).detach()
    logits = msgdecoder(x)
    decoded = torch.argmax(logits, dim=-1)
    acc = ((msg - decoded) == 0).float().mean()
    labels = F.one_hot(msg.long(), num_classes=2).float()
    loss = losses.binary_cross_entropy_with_logits(logits.float(), labels)
    loss.backward()
    return loss.detach(), acc.detach()
