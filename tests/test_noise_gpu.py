"""Hot path (ii), distortion stack on a B200: csrc/noise.cu through the drop-in `aqualora_b200.noise_layers` classes (ctypes ->
C ABI) against the reference's own JpegCompression outputs (tests/golden/jpeg_small.pt) and the CPU oracle
(oracle/noise_oracle.py) on identical explicit parameters.

Tolerance (fp32 kernels vs fp32 CPU, different summation order): 2e-5 absolute for the JPEG mask (64-term DCT sums of
values in [-3, 3]; same bound the oracle is pinned to the reference with), 1e-5 for the other layers on images in [-1, 1].
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _img(B, H, W, seed):
    return torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(seed)) * 2 - 1


@pytest.fixture(scope="module")
def nl(cuda_device):
    from aqualora_b200 import noise_layers

    return noise_layers


def test_jpeg_matches_reference_golden(nl, cuda_device, golden_dir):
    for c in torch.load(os.path.join(golden_dir, "jpeg_small.pt"), weights_only=False):   # 64x64, 40x72, 37x50 (pad / un-pad)
        got = nl.jpeg_mask(c["x"].to(cuda_device))
        torch.testing.assert_close(got.cpu(), c["y"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("B,H,W", [(2, 512, 512), (1, 8, 8), (1, 520, 1030), (3, 100, 36), (1, 511, 513)])
def test_jpeg_matches_oracle(nl, cuda_device, B, H, W):
    from oracle import noise_oracle as NO

    x = _img(B, H, W, H + W)
    got = nl.JpegCompression(cuda_device)([x.to(cuda_device), None])[0]
    torch.testing.assert_close(got.cpu(), NO.jpeg_mask(x), rtol=0, atol=2e-5)


def test_crop_resize_matches_oracle(nl, cuda_device):
    from oracle import noise_oracle as NO

    x = _img(2, 512, 512, 1)
    rng = np.random.default_rng(3)
    for _ in range(4):
        p = NO.draw_params(rng, 2, 2)
        got = nl.crop_resize(x.to(cuda_device), **p)
        torch.testing.assert_close(got.cpu(), NO.crop_resize(x, **p), rtol=1e-5, atol=1e-5)
    # extreme boxes: smallest crop to the largest resize, full-size crop
    for p in (dict(top=0, left=0, crop_h=256, crop_w=256, resize_h=511, resize_w=511),
              dict(top=1, left=1, crop_h=511, crop_w=511, resize_h=256, resize_w=300)):
        torch.testing.assert_close(nl.crop_resize(x.to(cuda_device), **p).cpu(), NO.crop_resize(x, **p), rtol=1e-5, atol=1e-5)
    # the layer class draws a box inside the image and records it
    layer = nl.CropandResize((256, 512), (256, 512), rng=np.random.default_rng(0))
    out = layer([x.to(cuda_device), None])[0]
    assert tuple(out.shape) == (2, 3, 512, 512)
    torch.testing.assert_close(out.cpu(), NO.crop_resize(x, **layer.last_params), rtol=1e-5, atol=1e-5)
    # a constant image stays constant
    c = torch.full((1, 3, 512, 512), 0.25, device=cuda_device)
    assert torch.allclose(nl.crop_resize(c, 10, 20, 300, 400, 260, 500), c, atol=1e-6)


def test_gaussian_blur_matches_oracle(nl, cuda_device):
    from oracle import noise_oracle as NO

    x = _img(3, 512, 512, 2)
    sig = [0.05, 3.3, 9.9]
    got = nl.gaussian_blur(x.to(cuda_device), sig)
    torch.testing.assert_close(got.cpu(), NO.gaussian_blur(x, sig), rtol=1e-5, atol=1e-5)
    x2 = _img(1, 37, 150, 5)                    # ragged tile edges, reflect border on every side
    torch.testing.assert_close(nl.gaussian_blur(x2.to(cuda_device), [2.0]).cpu(), NO.gaussian_blur(x2, [2.0]), rtol=1e-5, atol=1e-5)
    c = torch.full((1, 3, 64, 64), -0.5, device=cuda_device)
    assert torch.allclose(nl.gaussian_blur(c, [4.0]), c, atol=1e-6)      # normalised taps
    layer = nl.GaussianBlur(10.0, rng=np.random.default_rng(1))
    out = layer([x.to(cuda_device), None])[0]
    torch.testing.assert_close(out.cpu(), NO.gaussian_blur(x, layer.last_params["sigmas"]), rtol=1e-5, atol=1e-5)


def test_gaussian_noise_stream_and_oracle(nl, cuda_device):
    from oracle import noise_oracle as NO

    x = _img(2, 512, 512, 3)
    noise = nl.unit_noise(tuple(x.shape), seed=77, offset=5, device=cuda_device)
    got = nl.gaussian_noise(x.to(cuda_device), 0.13, seed=77, offset=5)
    torch.testing.assert_close(got.cpu(), NO.gaussian_noise(x, 0.13, noise.cpu()), rtol=1e-6, atol=1e-6)
    # counter-based stream: deterministic, offset k skips exactly 4k normals, seeds decorrelate
    again = nl.unit_noise(tuple(x.shape), seed=77, offset=5, device=cuda_device)
    assert torch.equal(noise, again)
    base = nl.unit_noise((4096,), seed=77, offset=0, device=cuda_device)
    assert torch.equal(base[20:], nl.unit_noise((4096 - 20,), seed=77, offset=5, device=cuda_device))
    other = nl.unit_noise((4096,), seed=78, offset=0, device=cuda_device)
    assert abs(torch.corrcoef(torch.stack([base, other]))[0, 1].item()) < 0.1
    big = nl.unit_noise((16, 3, 512, 512), seed=1, device=cuda_device)     # 12.6 M normals
    assert abs(big.mean().item()) < 2e-3 and abs(big.std().item() - 1) < 2e-3
    assert abs((big ** 4).mean().item() - 3) < 2e-2                       # kurtosis of a normal
    assert torch.isfinite(big).all()
    odd = nl.unit_noise((1001,), seed=77, device=cuda_device)             # n % 4 != 0 tail
    assert torch.equal(odd, base[:1001])


def test_color_jiggle_matches_oracle(nl, cuda_device):
    from oracle import noise_oracle as NO

    x = _img(3, 256, 256, 4)
    rng = np.random.default_rng(9)
    for _ in range(6):                                   # different op orders
        p = NO.draw_params(rng, 5, 3)
        got = nl.color_jiggle(x.to(cuda_device), **p)
        want = NO.color_jiggle(x, **p)
        # hue wraps: compare away from the HSV sector boundaries where a 1-ulp difference picks another branch
        diff = (got.cpu() - want).abs()
        assert (diff > 1e-4).float().mean().item() < 1e-5, p
        assert diff.median().item() < 1e-6
    layer = nl.ColorJitter(rng=np.random.default_rng(2))
    out = layer([x.to(cuda_device), None])[0]
    diff = (out.cpu() - NO.color_jiggle(x, **layer.last_params)).abs()
    assert (diff > 1e-4).float().mean().item() < 1e-5
    assert out.min().item() >= -1 - 1e-6 and out.max().item() <= 1 + 1e-6


def test_noiser_selection_follows_reference(nl, cuda_device):
    """noiser.py:41-44: one layer per call, drawn with the given probabilities; Identity occupies slot 0."""
    from oracle import noise_oracle as NO

    names = ["Jpeg", "CropandResize", "GaussianBlur", "GaussianNoise", "ColorJitter"]
    p = [0.4, 0.1, 0.2, 0.05, 0.1, 0.15]                       # train/latent_wm_pretrain.py:188 after epoch 12
    noiser = nl.Noiser(names, p, cuda_device, rng=np.random.default_rng(7))
    assert [type(l).__name__ for l in noiser.noise_layers] == ["Identity", "JpegCompression", "CropandResize", "GaussianBlur",
                                                               "GaussianNoise", "ColorJitter"]
    x = _img(2, 512, 512, 6).to(cuda_device)
    ref_rng = np.random.default_rng(7)
    seen = set()
    for _ in range(12):
        out = noiser([x.clone(), None])
        assert out[1] is None and tuple(out[0].shape) == (2, 3, 512, 512)
        want_idx = NO.draw_layer(ref_rng, p)
        assert noiser.last_layer == want_idx
        seen.add(want_idx)
        layer = noiser.noise_layers[want_idx]
        if want_idx == 0:
            assert torch.equal(out[0], x)
        elif want_idx == 1:
            torch.testing.assert_close(out[0].cpu(), NO.jpeg_mask(x.cpu()), rtol=0, atol=2e-5)
        # keep the reference generator in lock-step with what the layer consumed
        if want_idx == 2:
            for _k in range(4):
                ref_rng.integers(256, 512)
            ref_rng.integers(0, 512 - layer.last_params["crop_h"] + 1); ref_rng.integers(0, 512 - layer.last_params["crop_w"] + 1)
        elif want_idx == 3:
            for _k in range(2):
                ref_rng.random()
        elif want_idx == 4:
            ref_rng.random(); ref_rng.integers(0, 2 ** 62)
        elif want_idx == 5:
            for _k in range(8):
                ref_rng.random()
            ref_rng.permutation(4)
    assert len(seen) >= 3
    with pytest.raises(ValueError):
        nl.Noiser(["Rotation"], [0.5, 0.5], cuda_device)
