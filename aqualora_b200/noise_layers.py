"""Drop-in counterpart of the reference's `utils/noise_layers` package (noiser.py, noises.py, jpeg_compression.py,
identity.py): same class names, constructor signatures and calling convention (`layer([image, cover]) -> [image, cover]`),
backed by the sm_100a kernels in csrc/noise.cu.  There is no PyTorch fallback: CPU tensors raise.

The reference layers sample their parameters from numpy's / kornia's global RNG inside `forward`.  Here every layer draws
them on the host from an explicit `numpy.random.Generator` (`rng=`; default: a generator seeded from numpy's global RNG at
construction, so `np.random.seed` still controls a run) and hands them to the kernel as arguments, which is what makes the
CUDA path checkable against the CPU oracle on identical parameters.  `layer.last_params` records what was drawn.

    Noiser(noise_layers, posibilities, device)       noiser.py:12-44
    JpegCompression(device)                          jpeg_compression.py:67-162
    CropandResize(crop_range, resize_range)          noises.py:34-57
    GaussianBlur(blur)                               noises.py:59-70
    GaussianNoise(std)                               noises.py:72-85
    ColorJitter()                                    noises.py:88-104
    Identity()                                       identity.py
    distorsion_unit(image, type)                     noiser.py:46-71 (stage-3 parameter sets, same kernels)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops


def _default_rng():
    return np.random.default_rng(np.random.randint(0, 2 ** 31 - 1))


def random_float(rng, lo, hi):
    return float(rng.random() * (hi - lo) + lo)          # noises.py:8-15


def random_int(rng, lo, hi):
    return int(rng.integers(lo, hi))                     # noises.py:17-18 (np.random.randint: hi exclusive)


# functional forms (explicit parameters), differentiable w.r.t. the image --------------------------------------------------
# train/latent_wm_pretrain.py:186-216 back-propagates through `noiser(...)` into the encoder, so every layer is an
# autograd.Function whose backward is the matching input-gradient kernel (csrc/noise.cu); parameters get no gradient.
class _JpegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.noise_jpeg(x)

    @staticmethod
    def backward(ctx, gy):
        return ops.noise_jpeg_bwd(gy.contiguous())


class _CropResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, top, left, crop_h, crop_w, resize_h, resize_w, out_hw):
        ctx.box = (int(top), int(left), int(crop_h), int(crop_w), int(resize_h), int(resize_w))
        ctx.in_hw = tuple(x.shape[2:])
        return ops.noise_crop_resize(x, top, left, crop_h, crop_w, resize_h, resize_w, out_hw)

    @staticmethod
    def backward(ctx, gy):
        return (ops.noise_crop_resize_bwd(gy.contiguous(), ctx.in_hw, *ctx.box),) + (None,) * 7


class _BlurFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sigmas, ksize):
        ctx.save_for_backward(sigmas)
        ctx.ksize = ksize
        return ops.noise_gauss_blur(x, sigmas, ksize)

    @staticmethod
    def backward(ctx, gy):
        (sigmas,) = ctx.saved_tensors
        return ops.noise_gauss_blur_bwd(gy.contiguous(), sigmas, ctx.ksize), None, None


class _NoiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, std, seed, offset):
        return ops.noise_gauss_noise(x, std, seed, offset)

    @staticmethod
    def backward(ctx, gy):
        return gy, None, None, None            # y = x + std * n: identity


class _JiggleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, order):
        ctx.save_for_backward(x, params)
        ctx.order = tuple(order)
        return ops.noise_color_jiggle(x, params, order)

    @staticmethod
    def backward(ctx, gy):
        x, params = ctx.saved_tensors
        return ops.noise_color_jiggle_bwd(x, gy.contiguous(), params, ctx.order), None, None


def jpeg_mask(x: torch.Tensor) -> torch.Tensor:
    return _JpegFn.apply(x)


def crop_resize(x, top, left, crop_h, crop_w, resize_h, resize_w, out_hw=(512, 512)):
    return _CropResizeFn.apply(x, top, left, crop_h, crop_w, resize_h, resize_w, tuple(out_hw))


def gaussian_blur(x, sigmas, ksize=(3, 9)):
    s = torch.as_tensor(sigmas, dtype=torch.float32).to(x.device)
    return _BlurFn.apply(x, s, tuple(ksize))


def gaussian_noise(x, std, seed, offset=0):
    return _NoiseFn.apply(x, std, seed, offset)


def unit_noise(shape, seed, offset=0, device="cuda"):
    """The N(0, 1) tensor `gaussian_noise(x, std, seed, offset)` adds (times std) -- for parity tests against the oracle."""
    return ops.noise_gauss_noise(None, 1.0, seed, offset, shape=shape, device=device)


def color_jiggle(x, brightness, contrast, saturation, hue, order):
    p = torch.tensor([list(map(float, brightness)), list(map(float, contrast)), list(map(float, saturation)), list(map(float, hue))],
                     dtype=torch.float32).t().contiguous().to(x.device)
    return _JiggleFn.apply(x, p, tuple(int(o) for o in order))


# layer classes ------------------------------------------------------------------------------------------------------
class Identity(nn.Module):
    def forward(self, noised_and_cover):
        return noised_and_cover


class JpegCompression(nn.Module):
    def __init__(self, device=None, yuv_keep_weights=(25, 9, 9)):
        super().__init__()
        if tuple(yuv_keep_weights) != (25, 9, 9):
            raise ValueError("JpegCompression: the kernel is specialised for the reference's yuv_keep_weights = (25, 9, 9)")
        self.last_params = {}

    def forward(self, noised_and_cover):
        noised_and_cover[0] = jpeg_mask(noised_and_cover[0])
        return noised_and_cover


class CropandResize(nn.Module):
    def __init__(self, crop_size_range, resize_size_range, rng=None):
        super().__init__()
        self.crop_size_min, self.crop_size_max = crop_size_range
        self.resize_size_min, self.resize_size_max = resize_size_range
        self.rng = rng or _default_rng()
        self.last_params = {}

    def forward(self, noised_and_cover):
        x = noised_and_cover[0]
        H, W = x.shape[2:]
        ch = random_int(self.rng, self.crop_size_min, self.crop_size_max)
        cw = random_int(self.rng, self.crop_size_min, self.crop_size_max)
        rh = random_int(self.rng, self.resize_size_min, self.resize_size_max)
        rw = random_int(self.rng, self.resize_size_min, self.resize_size_max)
        if ch > H or cw > W:
            raise ValueError(f"Required crop size {(ch, cw)} is larger than input image size {(H, W)}")   # T.RandomCrop's error
        top = int(self.rng.integers(0, H - ch + 1))
        left = int(self.rng.integers(0, W - cw + 1))
        self.last_params = dict(top=top, left=left, crop_h=ch, crop_w=cw, resize_h=rh, resize_w=rw)
        noised_and_cover[0] = crop_resize(x, **self.last_params)
        return noised_and_cover


class GaussianBlur(nn.Module):
    def __init__(self, blur=2.0, rng=None, kernel_size=(3, 9), sigma_min=0.0):
        super().__init__()
        self.gaussian_blur_max = blur
        self.sigma_min = sigma_min
        self.kernel_size = kernel_size
        self.rng = rng or _default_rng()
        self.last_params = {}

    def forward(self, noised_and_cover):
        x = noised_and_cover[0]
        sig = [max(random_float(self.rng, self.sigma_min, self.gaussian_blur_max), 1e-3) for _ in range(x.shape[0])]
        self.last_params = dict(sigmas=sig)
        noised_and_cover[0] = gaussian_blur(x, sig, self.kernel_size)
        return noised_and_cover


class GaussianNoise(nn.Module):
    def __init__(self, std=0.1, rng=None):
        super().__init__()
        self.gaussian_std_max = std
        self.rng = rng or _default_rng()
        self.last_params = {}

    def forward(self, noised_and_cover):
        std = random_float(self.rng, 0, self.gaussian_std_max)
        seed = int(self.rng.integers(0, 2 ** 62))
        self.last_params = dict(std=std, seed=seed, offset=0)
        noised_and_cover[0] = gaussian_noise(noised_and_cover[0], std, seed, 0)
        return noised_and_cover


class ColorJitter(nn.Module):
    def __init__(self, rng=None, brightness=(0.7, 1.3), contrast=(0.8, 1.25), saturation=(0.8, 1.25), hue=(-0.2, 0.2)):
        super().__init__()
        self.ranges = (brightness, contrast, saturation, hue)
        self.rng = rng or _default_rng()
        self.last_params = {}

    def forward(self, noised_and_cover):
        x = noised_and_cover[0]
        B = x.shape[0]
        draw = lambda r: [random_float(self.rng, r[0], r[1]) for _ in range(B)]
        self.last_params = dict(brightness=draw(self.ranges[0]), contrast=draw(self.ranges[1]), saturation=draw(self.ranges[2]),
                                hue=draw(self.ranges[3]), order=[int(i) for i in self.rng.permutation(4)])
        noised_and_cover[0] = color_jiggle(x, **self.last_params)
        return noised_and_cover


class Noiser(nn.Module):
    """Picks ONE layer per call with `np.random.choice(p=posibilities)` and applies it to the whole batch (noiser.py:41-44).
    `noise_layers` is the reference's list of placeholder strings (or layer instances); Identity is always slot 0."""

    def __init__(self, noise_layers: list, posibilities: list, device=None, rng=None):
        super().__init__()
        self.rng = rng or _default_rng()
        self.noise_layers = [Identity()]
        for layer in noise_layers:
            if type(layer) is str:
                if layer == "Identity":
                    continue
                elif layer == "Jpeg":
                    self.noise_layers.append(JpegCompression(device))
                elif layer == "CropandResize":
                    self.noise_layers.append(CropandResize((256, 512), (256, 512), rng=self.rng))
                elif layer == "GaussianBlur":
                    self.noise_layers.append(GaussianBlur(10.0, rng=self.rng))
                elif layer == "GaussianNoise":
                    self.noise_layers.append(GaussianNoise(0.2, rng=self.rng))
                elif layer == "ColorJitter":
                    self.noise_layers.append(ColorJitter(rng=self.rng))
                else:
                    raise ValueError("Wrong layer placeholder string in Noiser.__init__().")
            else:
                self.noise_layers.append(layer)
        self.posibilities = posibilities
        self.last_layer = None

    def forward(self, encoded_and_cover, possibilites=None):
        p = self.posibilities if possibilites is None else possibilites
        idx = int(self.rng.choice(len(self.noise_layers), p=np.asarray(p, dtype=np.float64)))
        self.last_layer = idx
        return self.noise_layers[idx](encoded_and_cover)


def distorsion_unit(encoded_image, type, rng=None):
    """Stage-3 (robustness fine-tuning) distortions, noiser.py:46-71: milder parameter sets on the same kernels.  Images are in
    [0, 1] here (the reference applies these to pipeline outputs before normalisation)."""
    rng = rng or _default_rng()
    B, _, H, W = encoded_image.shape
    if type == "color_jitter":
        layer = ColorJitter(rng, (0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
        return (layer([encoded_image * 2 - 1, None])[0] + 1) / 2      # the kernel maps [-1, 1] <-> [0, 1] around the jiggle
    if type == "crop":
        ch, cw = random_int(rng, 432, 512), random_int(rng, 432, 512)
        top, left = int(rng.integers(0, H - ch + 1)), int(rng.integers(0, W - cw + 1))
        return crop_resize(encoded_image, top, left, ch, cw, 512, 512, (512, 512))   # second resize is the identity
    if type == "blur":
        return gaussian_blur(encoded_image, [4.0] * B, (3, 5))
    if type == "noise":
        return gaussian_noise(encoded_image, 0.1, int(rng.integers(0, 2 ** 62)), 0).clamp(0, 1)
    raise ValueError("Wrong distorsion type.")
