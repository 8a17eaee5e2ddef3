"""Second-source checks for the oracle pieces whose first source (kornia 0.6.12) is neither installed nor vendored
("parity unpinned" in DESIGN.md section 2): where an independent implementation of the SAME arithmetic exists in this image, the
restatement is held to it.  This does not pin kornia's parameter conventions (kernel-size order, additive brightness, multiplicative
contrast) -- those stay restated from its published behaviour -- but it does pin the numerics they are built from.

  * gaussian_blur  <-> torchvision.transforms.functional.gaussian_blur (same normalised exp(-x^2 / 2 sigma^2) taps, reflect border)
  * rgb_to_hsv / hsv_to_rgb <-> the standard library's colorsys
  * the hue rotation of color_jiggle <-> torchvision.transforms.functional.adjust_hue
"""
import colorsys
import math

import pytest
import torch

from oracle import noise_oracle as NO


@pytest.mark.parametrize("sigma", [0.3, 1.0, 4.5, 10.0])
def test_gaussian_blur_matches_torchvision(sigma):
    from torchvision.transforms import functional as TF

    x = torch.rand(2, 3, 40, 56, generator=torch.Generator().manual_seed(3)) * 2 - 1
    got = NO.gaussian_blur(x, [sigma, sigma], ksize=(3, 9))                      # (ky, kx)
    want = TF.gaussian_blur(x, kernel_size=[9, 3], sigma=[sigma, sigma])         # torchvision: [kx, ky]
    torch.testing.assert_close(got, want, rtol=0, atol=2e-6)


def test_hsv_round_trip_matches_colorsys():
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(1, 3, 16, 16, generator=g)
    rgb[0, :, 0, 0] = 0.0                                                       # black, white, grey and a pure primary
    rgb[0, :, 0, 1] = 1.0
    rgb[0, :, 0, 2] = 0.4
    rgb[0, :, 0, 3] = torch.tensor([1.0, 0.0, 0.0])
    hsv = NO.rgb_to_hsv(rgb)
    back = NO.hsv_to_rgb(hsv)
    torch.testing.assert_close(back, rgb, rtol=0, atol=1e-6)
    for y in range(16):
        for x in range(16):
            r, gch, b = (float(v) for v in rgb[0, :, y, x])
            h, s, v = colorsys.rgb_to_hsv(r, gch, b)
            assert float(hsv[0, 0, y, x]) / (2 * math.pi) == pytest.approx(h, abs=1e-5)
            assert float(hsv[0, 1, y, x]) == pytest.approx(s, abs=1e-5)
            assert float(hsv[0, 2, y, x]) == pytest.approx(v, abs=1e-6)


@pytest.mark.parametrize("hue", [-0.3, -0.05, 0.0, 0.1, 0.45])
def test_hue_rotation_matches_torchvision(hue):
    from torchvision.transforms import functional as TF

    x = torch.rand(2, 3, 24, 24, generator=torch.Generator().manual_seed(5)) * 2 - 1
    got = NO.color_jiggle(x, [1.0, 1.0], [1.0, 1.0], [1.0, 1.0], [hue, hue], order=(3,))     # hue only
    want = TF.adjust_hue(x / 2 + 0.5, hue) * 2 - 1
    torch.testing.assert_close(got, want, rtol=0, atol=2e-5)


def test_identity_parameters_leave_the_image_unchanged():
    x = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(6)) * 2 - 1
    got = NO.color_jiggle(x, [1.0, 1.0], [1.0, 1.0], [1.0, 1.0], [0.0, 0.0], order=(0, 1, 2, 3))
    torch.testing.assert_close(got, x, rtol=0, atol=2e-6)
    torch.testing.assert_close(NO.gaussian_noise(x, 0.0, torch.randn_like(x)), x, rtol=0, atol=0)
