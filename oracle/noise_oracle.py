"""CPU oracle for hot path (ii), distortion part: utils/noise_layers.  TEST INFRASTRUCTURE ONLY.

Every random quantity the reference draws inside a layer (crop box, sizes, sigma, jitter factors, op order, the noise
tensor) is an EXPLICIT argument here, so the oracle and the CUDA kernels consume identical parameters.

Parity status
  * jpeg_mask: PINNED against the reference's own JpegCompression (tests/golden/jpeg_small.pt).
  * crop_resize: torchvision `T.RandomCrop` + `T.Resize(antialias=None)` on tensors == slicing + F.interpolate(bilinear,
    align_corners=False, antialias=False); restated with torch ops (torchvision is installed, behaviour identical).
  * gaussian_blur / gaussian_noise / color_jiggle follow kornia==0.6.12 (requirements.txt:13), which is NOT installed
    here and not vendored: its published semantics are restated from memory and are "parity unpinned" (DESIGN.md):
    RandomGaussianBlur((ky, kx), sigma range) -> separable normalised Gaussian taps, reflect border, one sigma per
    sample for both axes; ColorJiggle -> brightness x + (b - 1), contrast x * c, saturation in HSV, hue shift h * 2pi in
    HSV, each clamped to [0, 1], applied in a sampled order.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# JPEG-mask  (utils/noise_layers/jpeg_compression.py)
# ----------------------------------------------------------------------------------------------


def zigzag_keep_mask(keep_count: int, window: int = 8) -> torch.Tensor:
    """jpeg_compression.py:31-41: first `keep_count` positions of the 8x8 zig-zag order."""
    order = sorted(((x, y) for x in range(window) for y in range(window)),
                   key=lambda p: (p[0] + p[1], -p[1] if (p[0] + p[1]) % 2 else p[1]))
    m = torch.zeros(window, window)
    for i, j in order[:keep_count]:
        m[i, j] = 1
    return m


def dct_matrices(dtype=torch.float32):
    """1-D factors of the reference's separable 64-tap filters (jpeg_compression.py:8-18,44-50):
    D[k, n] = cos(pi/8 (n + 1/2) k)  and  I[n, k] = ((k == 0 ? -1/2 : 0) + cos(pi/8 (n + 1/2) k)) * sqrt(1/16)."""
    n = torch.arange(8, dtype=torch.float64)
    k = torch.arange(8, dtype=torch.float64)
    d = torch.cos(math.pi / 8 * (n[None, :] + 0.5) * k[:, None])                       # [k, n]
    i = ((k[None, :] == 0).double() * (-0.5) + torch.cos(math.pi / 8 * (n[:, None] + 0.5) * k[None, :])) * math.sqrt(1 / 16)
    return d.to(dtype), i.to(dtype)


RGB2YUV = ((0.299, 0.587, 0.114), (-0.14713, -0.28886, 0.436), (0.615, -0.51499, -0.10001))   # jpeg_compression.py:53-57
YUV2RGB = ((1.0, 0.0, 1.13983), (1.0, -0.39465, -0.58060), (1.0, 2.03211, 0.0))               # jpeg_compression.py:60-64


def jpeg_mask(x: torch.Tensor, keep=(25, 9, 9)) -> torch.Tensor:
    """JpegCompression.forward (jpeg_compression.py:130-162): zero-pad to a multiple of 8, RGB->YUV, 8x8 block DCT, keep
    the first (25, 9, 9) zig-zag coefficients of (Y, U, V), inverse DCT, YUV->RGB, un-pad.  No quantisation."""
    B, C, H, W = x.shape
    ph, pw = (8 - H % 8) % 8, (8 - W % 8) % 8
    xp = F.pad(x, (0, pw, 0, ph))
    m = torch.tensor(RGB2YUV, dtype=x.dtype, device=x.device)
    yuv = torch.einsum("oc,bchw->bohw", m, xp)
    Hp, Wp = xp.shape[2:]
    blocks = yuv.view(B, 3, Hp // 8, 8, Wp // 8, 8).permute(0, 1, 2, 4, 3, 5)           # [B, 3, by, bx, y, x]
    d, i = (t.to(x.device) for t in dct_matrices(x.dtype))
    coef = torch.einsum("ky,...yx,lx->...kl", d, blocks, d)                               # [.., ky, kx]
    masks = torch.stack([zigzag_keep_mask(k) for k in keep]).to(device=x.device, dtype=x.dtype)   # [3, ky, kx]
    coef = coef * masks[None, :, None, None]
    rec = torch.einsum("yk,...kl,xl->...yx", i, coef, i)
    rec = rec.permute(0, 1, 2, 4, 3, 5).reshape(B, 3, Hp, Wp)
    rgb = torch.einsum("oc,bchw->bohw", torch.tensor(YUV2RGB, dtype=x.dtype, device=x.device), rec)
    return rgb[:, :, :H, :W].clone()


# ----------------------------------------------------------------------------------------------
# CropandResize  (utils/noise_layers/noises.py:34-57)
# ----------------------------------------------------------------------------------------------
def crop_resize(x: torch.Tensor, top: int, left: int, crop_h: int, crop_w: int, resize_h: int, resize_w: int,
                out_hw=(512, 512)) -> torch.Tensor:
    """T.RandomCrop((crop_h, crop_w)) at (top, left) -> T.Resize((resize_h, resize_w), antialias=None) ->
    T.Resize(out_hw, antialias=None); one box for the whole batch."""
    y = x[:, :, top:top + crop_h, left:left + crop_w]
    y = F.interpolate(y, size=(resize_h, resize_w), mode="bilinear", align_corners=False, antialias=False)
    return F.interpolate(y, size=out_hw, mode="bilinear", align_corners=False, antialias=False)


# ----------------------------------------------------------------------------------------------
# GaussianBlur  (noises.py:59-70 -> kornia RandomGaussianBlur((3, 9), (0, 10), p=1))
# ----------------------------------------------------------------------------------------------
def gaussian_taps(ksize: int, sigma: float) -> torch.Tensor:
    xs = torch.arange(ksize, dtype=torch.float32) - ksize // 2
    g = torch.exp(-(xs ** 2) / (2.0 * sigma * sigma))
    return g / g.sum()


def gaussian_blur(x: torch.Tensor, sigmas, ksize=(3, 9)) -> torch.Tensor:
    """Per-sample sigma (same for both axes); kernel (ky, kx) = (3, 9); reflect border; separable."""
    ky, kx = ksize
    out = torch.empty_like(x)
    for b in range(x.shape[0]):
        s = float(sigmas[b])
        ty, tx = gaussian_taps(ky, s).to(x.device), gaussian_taps(kx, s).to(x.device)
        xb = F.pad(x[b:b + 1], (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect")
        C = x.shape[1]
        xb = F.conv2d(xb, tx.view(1, 1, 1, kx).repeat(C, 1, 1, 1), groups=C)
        out[b:b + 1] = F.conv2d(xb, ty.view(1, 1, ky, 1).repeat(C, 1, 1, 1), groups=C)
    return out


# ----------------------------------------------------------------------------------------------
# GaussianNoise  (noises.py:72-85)
# ----------------------------------------------------------------------------------------------
def gaussian_noise(x: torch.Tensor, std: float, noise: torch.Tensor) -> torch.Tensor:
    """x + std * N(0, 1) with the unit noise tensor given explicitly."""
    return x + std * noise


# ----------------------------------------------------------------------------------------------
# ColorJitter  (noises.py:88-104 -> kornia ColorJiggle)
# ----------------------------------------------------------------------------------------------
def rgb_to_hsv(rgb: torch.Tensor, eps: float = 1e-8):
    """kornia.color.rgb_to_hsv: h in [0, 2pi), s, v in [0, 1]."""
    r, g, b = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    maxc, _ = rgb.max(dim=1)
    minc, _ = rgb.min(dim=1)
    v = maxc
    delta = maxc - minc
    s = delta / (maxc + eps)
    dz = torch.where(delta == 0, torch.ones_like(delta), delta)
    rc, gc, bc = (maxc - r), (maxc - g), (maxc - b)
    h = torch.where(maxc == r, bc - gc, torch.where(maxc == g, 2.0 * dz + rc - bc, 4.0 * dz + gc - rc))
    h = (h / dz / 6.0) % 1.0
    return torch.stack([h * 2 * math.pi, s, v], dim=1)


def hsv_to_rgb(hsv: torch.Tensor):
    h = hsv[:, 0] / (2 * math.pi)
    s, v = hsv[:, 1], hsv[:, 2]
    hi = torch.floor(h * 6) % 6
    f = (h * 6) % 6 - hi
    p = v * (1 - s)
    q = v * (1 - f * s)
    t = v * (1 - (1 - f) * s)
    hi = hi.long()
    r = torch.stack([v, q, p, p, t, v], dim=0).gather(0, hi[None])[0]
    g = torch.stack([t, v, v, q, p, p], dim=0).gather(0, hi[None])[0]
    b = torch.stack([p, p, t, v, v, q], dim=0).gather(0, hi[None])[0]
    return torch.stack([r, g, b], dim=1)


def color_jiggle(x: torch.Tensor, brightness, contrast, saturation, hue, order) -> torch.Tensor:
    """noises.py:95-104: x in [-1, 1] -> [0, 1], ColorJiggle with per-sample factors and one op order, back to [-1, 1].
    order is a permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue)."""
    img = x / 2 + 0.5
    b = torch.as_tensor(brightness, dtype=x.dtype, device=x.device).view(-1, 1, 1, 1)
    c = torch.as_tensor(contrast, dtype=x.dtype, device=x.device).view(-1, 1, 1, 1)
    s = torch.as_tensor(saturation, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    h = torch.as_tensor(hue, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    for op in order:
        if op == 0:
            img = (img + (b - 1)).clamp(0, 1)
        elif op == 1:
            img = (img * c).clamp(0, 1)
        elif op == 2:
            hsv = rgb_to_hsv(img)
            hsv = torch.stack([hsv[:, 0], (hsv[:, 1] * s).clamp(0, 1), hsv[:, 2]], dim=1)
            img = hsv_to_rgb(hsv)
        else:
            hsv = rgb_to_hsv(img)
            hh = torch.fmod(hsv[:, 0] + h * 2 * math.pi, 2 * math.pi)
            img = hsv_to_rgb(torch.stack([hh, hsv[:, 1], hsv[:, 2]], dim=1))
    return img * 2 - 1


# ----------------------------------------------------------------------------------------------
# Noiser  (utils/noise_layers/noiser.py:12-44)
# ----------------------------------------------------------------------------------------------
LAYER_NAMES = ("Identity", "Jpeg", "CropandResize", "GaussianBlur", "GaussianNoise", "ColorJitter")


def draw_layer(rng: np.random.Generator, probabilities) -> int:
    """noiser.py:41-44: np.random.choice over the layer list with the given probabilities (one layer per call)."""
    return int(rng.choice(len(probabilities), p=np.asarray(probabilities, dtype=np.float64)))


def draw_params(rng: np.random.Generator, layer: int, batch: int, hw=(512, 512)) -> dict:
    """Sample the parameters each layer draws internally (noises.py:46-50, :67-68, :80-83, :96-102)."""
    H, W = hw
    if layer == 2:
        ch, cw = int(rng.integers(256, 512)), int(rng.integers(256, 512))
        rh, rw = int(rng.integers(256, 512)), int(rng.integers(256, 512))
        top, left = int(rng.integers(0, H - ch + 1)), int(rng.integers(0, W - cw + 1))
        return dict(top=top, left=left, crop_h=ch, crop_w=cw, resize_h=rh, resize_w=rw)
    if layer == 3:
        return dict(sigmas=[max(float(rng.uniform(0, 10.0)), 1e-3) for _ in range(batch)])
    if layer == 4:
        return dict(std=float(rng.uniform(0, 0.2)))
    if layer == 5:
        return dict(brightness=[float(rng.uniform(0.7, 1.3)) for _ in range(batch)],
                    contrast=[float(rng.uniform(0.8, 1.25)) for _ in range(batch)],
                    saturation=[float(rng.uniform(0.8, 1.25)) for _ in range(batch)],
                    hue=[float(rng.uniform(-0.2, 0.2)) for _ in range(batch)],
                    order=[int(i) for i in rng.permutation(4)])
    return {}


def apply_layer(x: torch.Tensor, layer: int, params: dict, noise: torch.Tensor | None = None) -> torch.Tensor:
    if layer == 0:
        return x
    if layer == 1:
        return jpeg_mask(x)
    if layer == 2:
        return crop_resize(x, **params)
    if layer == 3:
        return gaussian_blur(x, params["sigmas"])
    if layer == 4:
        return gaussian_noise(x, params["std"], noise)
    if layer == 5:
        return color_jiggle(x, params["brightness"], params["contrast"], params["saturation"], params["hue"], params["order"])
    raise ValueError(layer)
