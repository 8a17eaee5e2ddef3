"""CPU oracle for the caller of hot path (ii): one step of train/latent_wm_pretrain.py:164-217.  TEST INFRASTRUCTURE ONLY.

BASELINE.json configs[0] ("latent_wm_pretrain.py 48-bit msg, 64x64 VAE latents, batch=2 on CPU (plumbing, no GPU)") is exercised by
tests/test_pretrain_plumbing.py through this file.  Each function cites the reference lines it follows.

Parity status
  * prvl_loss, gen_combined_latents, draw_cornerfy: PINNED -- tests/golden/pretrain_small.pt holds outputs of the reference's own
    source for these functions (tools/gen_golden.py executes the function bodies read from train/latent_wm_pretrain.py; the script as
    a whole cannot be imported: accelerate / diffusers / lpips / torchsummary are absent).
  * SecretEncoderRef: same arithmetic as models_oracle.secret_encoder_forward (pinned to the reference's SecretEncoder) as an
    nn.Module with the reference's state-dict keys, so that it can be trained and checkpointed.
  * SecretDecoderRef: the reference's decoder IS torchvision's efficientnet_b1 with a replaced classifier (utils/models.py:84-96).
  * StubVAE, lpips_stub: stand-ins for the third-party frozen networks (AutoencoderKL, LPIPS-VGG: out of scope, no weights offline).
    They only give the step the right tensor shapes.  "parity unpinned" for those two, by construction.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import noise_oracle as NO

WINDOW_SIZE = 32   # train/latent_wm_pretrain.py:39


def prvl_loss(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """train/latent_wm_pretrain.py:42-50: max over all positions of the 32x32 box mean of the channel-mean absolute difference
    (zero padding 16: a 513 x 513 map for 512 x 512 images)."""
    kernel = torch.ones((1, 1, WINDOW_SIZE, WINDOW_SIZE), dtype=torch.float32, device=img1.device) / (WINDOW_SIZE ** 2)
    diff = torch.abs(img1 - img2).mean(dim=1, keepdim=True)
    return F.conv2d(diff, kernel, padding=WINDOW_SIZE // 2).max()


def draw_cornerfy(rng: random.Random):
    """The random draws of gen_combined_latents in the reference's order (train/latent_wm_pretrain.py:134-137):
    choice([True, False, False, False]), then two uniform(1, 2) only when cornerfy is on."""
    cornerfy = rng.choice([True, False, False, False])
    hs, ws = (rng.uniform(1.0, 2.0), rng.uniform(1.0, 2.0)) if cornerfy else (1.0, 1.0)
    return cornerfy, hs, ws


def gen_combined_latents(latents, wm_latent, scale=1.0, cornerfy=False, height_scale=1.0, width_scale=1.0):
    """train/latent_wm_pretrain.py:133-149 with the random choices passed in: when `cornerfy`, the four corner quadrants of the
    watermark residual are pasted into the corners of a zero canvas enlarged by (height_scale, width_scale), which is then resized
    back (bilinear) -- the residual gets squeezed towards the corners; latents + residual * scale."""
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


class _View(nn.Module):
    def __init__(self, *shape):
        super().__init__()
        self.shape = shape

    def forward(self, x):
        return x.view(*self.shape)


class _Repeat(nn.Module):
    def __init__(self, *sizes):
        super().__init__()
        self.sizes = sizes

    def forward(self, x):
        return x.repeat(1, *self.sizes)


class SecretEncoderRef(nn.Module):
    """utils/models.py:51-81 (state-dict keys secret_scaler.0.* and secret_scaler.5.*; the last conv is zero-initialised, :63)."""

    def __init__(self, secret_len, base_res=32, resolution=64):
        super().__init__()
        f = resolution // base_res
        conv = nn.Conv2d(4, 4, 3, padding=1)
        nn.init.zeros_(conv.weight)
        nn.init.zeros_(conv.bias)
        self.secret_scaler = nn.Sequential(nn.Linear(secret_len, base_res * base_res), nn.SiLU(), _View(-1, 1, base_res, base_res),
                                           _Repeat(4, 1, 1), nn.Upsample(scale_factor=(f, f)), conv)

    def forward(self, x, c):
        c = F.interpolate(self.secret_scaler(c), size=(x.shape[2], x.shape[3]), mode="bilinear")
        return x + c, c


class SecretDecoderRef(nn.Module):
    """utils/models.py:84-96 without the ImageNet download (weights=None)."""

    def __init__(self, output_size=48):
        super().__init__()
        from torchvision.models import efficientnet_b1

        self.output_size = output_size
        self.model = efficientnet_b1(weights=None)
        self.model.classifier[1] = nn.Linear(self.model.classifier[1].in_features, output_size * 2, bias=True)

    def forward(self, x):
        x = F.interpolate(x, size=(512, 512), mode="bilinear")
        return self.model(x).view(-1, self.output_size, 2)


class StubVAE(nn.Module):
    """Shape-only stand-in for AutoencoderKL (frozen, third party): encode = 8x average pool + fixed 3->4 mix, decode = bilinear x8 +
    fixed 4->3 mix + tanh.  Seeded, frozen."""

    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("enc_mix", torch.randn(4, 3, 1, 1, generator=g) * 0.5)
        self.register_buffer("dec_mix", torch.randn(3, 4, 1, 1, generator=g) * 0.5)

    def encode(self, image):
        return F.conv2d(F.avg_pool2d(image, 8), self.enc_mix)

    def decode(self, latents):
        return torch.tanh(F.conv2d(F.interpolate(latents, scale_factor=8, mode="bilinear"), self.dec_mix))


def lpips_stub(a, b):
    """Stand-in for lpips.LPIPS(net='vgg') (third party, no weights offline): mean squared difference."""
    return ((a - b) ** 2).mean()


def pretrain_step(enc, dec, vae, image, msg, layer: int, layer_params: dict, rng: random.Random, warmup: bool, stage: int,
                  noise: torch.Tensor | None = None):
    """Loop body of train/latent_wm_pretrain.py:173-217 up to `loss.backward()`.  `stage` selects the loss mix of :208-214
    (0: message loss only, 1: + lpips, 2: 5 lpips + msg + 1.5 prvl); `layer` / `layer_params` are the Noiser's draw
    (utils/noise_layers/noiser.py:41-44) made explicit."""
    latents = vae.encode(image).detach()                                                       # :176
    _, wm_latent = enc(latents, msg.float())                                                   # :179
    cornerfy, hs, ws = draw_cornerfy(rng)
    wl = gen_combined_latents(latents, wm_latent, 0.03 if warmup else 1.0, cornerfy, hs, ws)   # :181-184
    clean_image = vae.decode(latents).detach()                                                 # :185
    wm_image = vae.decode(wl)                                                                  # :186
    lp = lpips_stub(clean_image, wm_image)                                                     # :187 (stub)
    prvl = prvl_loss(clean_image, wm_image)                                                    # :188
    distorted = NO.apply_layer(wm_image, layer, layer_params, noise)                           # :190-193
    reveal = dec(distorted)                                                                    # :195
    labels = F.one_hot(msg.long(), num_classes=2).float()                                      # :198
    msgloss = F.binary_cross_entropy_with_logits(reveal, labels)                               # :200
    if warmup or stage == 0:
        loss = msgloss
    elif stage == 1:
        loss = lp + msgloss
    else:
        loss = lp * 5 + msgloss * 1.0 + prvl * 1.5
    loss.backward()                                                                            # :216
    return {"loss": loss.detach(), "msgloss": msgloss.detach(), "lpips": lp.detach(), "prvl": prvl.detach(), "reveal": reveal.detach(),
            "wm_image": wm_image.detach(), "cornerfy": cornerfy}


def checkpoint_dict(enc, dec):
    """train/latent_wm_pretrain.py:246-249 (read back at train/ppft_train.py:550-554)."""
    return {"sec_decoder": dec.state_dict(), "sec_encoder": enc.state_dict()}
