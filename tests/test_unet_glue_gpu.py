"""SURVEY.md 8(f2) on a B200: csrc/unet_norm_act.cu (GroupNorm [+ time-embedding add] [+ SiLU] on channels-last bf16 rows, GEGLU)
through the C ABI, against a plain PyTorch fp32 reference of the op sequence it replaces
(scripts/lib/original_unet.py:440-453, :826, :1416 and :708-729), forward and backward, on identical bf16 inputs.

Tolerance: the kernels compute in fp32 and round the OUTPUT to bf16 once, so an element may differ from the fp32 reference by
half a bf16 ulp (2^-9 relative) plus the rounding of sums in a different order: |got - want| <= 2^-8 |want| + 2e-3 max|want|... the
absolute term covers elements that cancel to ~0.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

REL = 2.0 ** -8


def _close(got, want, abs_frac=2e-3):
    got, want = got.float().cpu(), want.float().cpu()
    bound = REL * want.abs() + abs_frac * want.abs().max().clamp_min(1e-6)
    bad = (got - want).abs() > bound
    assert not bad.any(), f"{int(bad.sum())} of {bad.numel()} elements off; worst {(got - want).abs().max().item():.3e}"


@pytest.mark.parametrize("B,C,H,W,G,eps,silu,add", [
    (2, 320, 16, 16, 32, 1e-5, True, False),     # ResnetBlock2D.norm1 + silu, cpg 10: a 16-byte load straddles two groups
    (2, 320, 16, 16, 32, 1e-5, True, True),      # + time embedding -> norm2 + silu
    (3, 64, 5, 7, 8, 1e-6, False, False),        # Transformer2DModel.norm (no activation), odd H x W, ragged row slabs
    (2, 960, 8, 8, 32, 1e-5, True, True),        # concatenated skip input, cpg 30
    (1, 2560, 8, 8, 32, 1e-5, True, False),      # widest ResNet input of SD 1.5, one row per thread pass
    (2, 32, 4, 4, 8, 1e-5, True, True),          # the tiny test U-Net: cpg 4 (more than two groups per 16-byte load)
    (16, 320, 64, 64, 32, 1e-5, True, True),     # BASELINE size: 42 MB per tensor
])
def test_group_norm_nhwc_matches_fp32_reference(cuda_device, B, C, H, W, G, eps, silu, add):
    from aqualora_b200.unet_ops import group_norm_nhwc

    g = torch.Generator().manual_seed(C + H + int(silu))
    x = (torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(C, generator=g)).bfloat16()
    t = (0.5 * torch.randn(B, C, generator=g)).bfloat16() if add else None
    dy = torch.randn(B, C, H, W, generator=g).bfloat16()

    xr = x.float().requires_grad_(True)
    h = xr + t.float()[:, :, None, None] if add else xr
    want = F.group_norm(h, G, gamma.float(), beta.float(), eps)
    if silu:
        want = F.silu(want)
    want.backward(dy.float())

    xd = x.to(cuda_device).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = group_norm_nhwc(xd, gamma.to(cuda_device), beta.to(cuda_device), G, eps, silu, None if t is None else t.to(cuda_device))
    assert got.shape == x.shape and got.is_contiguous(memory_format=torch.channels_last) and got.dtype == torch.bfloat16
    got.backward(dy.to(cuda_device))
    _close(got.detach(), want.detach())
    _close(xd.grad, xr.grad)


def test_group_norm_accepts_nchw_and_rejects_trainable_affine(cuda_device):
    from aqualora_b200._lib import AqualoraError
    from aqualora_b200.unet_ops import group_norm_nhwc

    x = torch.randn(2, 64, 6, 6).bfloat16().to(cuda_device)            # NCHW strides: converted once, result channels_last
    gamma = torch.ones(64, dtype=torch.bfloat16, device=cuda_device)
    beta = torch.zeros(64, dtype=torch.bfloat16, device=cuda_device)
    y = group_norm_nhwc(x, gamma, beta, 8, 1e-5, False)
    _close(y, F.group_norm(x.float(), 8, eps=1e-5))
    with pytest.raises(AqualoraError):
        group_norm_nhwc(x, gamma.clone().requires_grad_(True), beta, 8, 1e-5, False)
    with pytest.raises(AqualoraError):
        group_norm_nhwc(x.cpu(), gamma.cpu(), beta.cpu(), 8, 1e-5, False)          # no CPU path
    with pytest.raises(AqualoraError):
        group_norm_nhwc(x[:, :60], gamma[:60], beta[:60], 8, 1e-5, False)           # C % G != 0


@pytest.mark.parametrize("lead,F_", [((2, 77), 1280), ((3, 5, 7), 64), ((16, 4096), 1280)])
def test_geglu_matches_fp32_reference(cuda_device, lead, F_):
    from aqualora_b200.unet_ops import geglu

    g = torch.Generator().manual_seed(F_)
    p = (torch.randn(*lead, 2 * F_, generator=g) * 1.5).bfloat16()
    go = torch.randn(*lead, F_, generator=g).bfloat16()
    pr = p.float().requires_grad_(True)
    h, gate = pr.chunk(2, dim=-1)
    want = h * F.gelu(gate)
    want.backward(go.float())
    pd = p.to(cuda_device).requires_grad_(True)
    got = geglu(pd)
    got.backward(go.to(cuda_device))
    assert got.shape == want.shape
    _close(got.detach(), want.detach(), abs_frac=1e-4)
    _close(pd.grad, pr.grad, abs_frac=1e-4)


@pytest.mark.parametrize("lead,C,eps", [((2, 77), 320, 1e-5), ((3, 5), 64, 1e-5), ((2, 9), 1280, 1e-6), ((5,), 2048, 1e-5),
                                        ((7,), 776, 1e-5), ((16, 4096), 320, 1e-5)])
def test_layer_norm_matches_fp32_reference(cuda_device, lead, C, eps):
    from aqualora_b200.unet_ops import layer_norm

    g = torch.Generator().manual_seed(C)
    x = (torch.randn(*lead, C, generator=g) * 2 + 0.5).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(C, generator=g)).bfloat16()
    dy = torch.randn(*lead, C, generator=g).bfloat16()
    xr = x.float().requires_grad_(True)
    want = F.layer_norm(xr, (C,), gamma.float(), beta.float(), eps)
    want.backward(dy.float())
    xd = x.to(cuda_device).requires_grad_(True)
    got = layer_norm(xd, gamma.to(cuda_device), beta.to(cuda_device), eps)
    got.backward(dy.to(cuda_device))
    _close(got.detach(), want.detach())
    _close(xd.grad, xr.grad)


@pytest.mark.parametrize("B,C,H,W", [(2, 320, 16, 16), (3, 64, 5, 7), (16, 640, 32, 32)])
def test_residual_add_bias_matches_fp32_reference(cuda_device, B, C, H, W):
    from aqualora_b200.unet_ops import residual_add_bias

    g = torch.Generator().manual_seed(C + H)
    a = torch.randn(B, C, H, W, generator=g).bfloat16()
    b = torch.randn(B, C, H, W, generator=g).bfloat16()
    bias = torch.randn(C, generator=g).bfloat16()
    cl = torch.channels_last
    ad = a.to(cuda_device).contiguous(memory_format=cl).requires_grad_(True)
    bd = b.to(cuda_device).requires_grad_(True)                      # NCHW strides: converted
    got = residual_add_bias(ad, bd, bias.to(cuda_device))
    want = a.float() + b.float() + bias.float()[None, :, None, None]
    _close(got.detach(), want, abs_frac=1e-6)
    dy = torch.randn(B, C, H, W, generator=g).bfloat16().to(cuda_device)
    got.backward(dy)
    assert torch.equal(ad.grad, dy) and torch.equal(bd.grad, dy)


def test_unet_forward_uses_glue_kernels_and_matches_library_ops(cuda_device):
    """The tiny U-Net with the glue kernels vs the same module tree with the library op sequence (trainable affine parameters
    switch the dispatch off): outputs agree to bf16 noise, and the fused run launches our kernels."""
    from aqualora_b200 import _lib, ppft
    from aqualora_b200.unet import UNetConfig

    cfg = UNetConfig.tiny(16)
    unet = ppft.build_unet(cfg, cuda_device, seed=3)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 16, 16, generator=g).bfloat16().to(cuda_device)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).bfloat16().to(cuda_device)
    t = torch.tensor([10, 500], device=cuda_device)
    n0 = _lib.load().aq_launch_count()
    with torch.no_grad():
        fused = unet(x, t, ctx).sample
    n_fused = _lib.load().aq_launch_count() - n0
    for m in unet.modules():
        if isinstance(m, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
            m.weight.requires_grad_(True)
    n0 = _lib.load().aq_launch_count()
    with torch.no_grad():
        plain = unet(x, t, ctx).sample
    n_plain = _lib.load().aq_launch_count() - n0
    n_gn = sum(isinstance(m, torch.nn.GroupNorm) for m in unet.modules())
    n_ln = sum(isinstance(m, torch.nn.LayerNorm) for m in unet.modules())
    from aqualora_b200.unet import ResnetBlock2D
    n_res = sum(isinstance(m, ResnetBlock2D) for m in unet.modules())
    assert n_fused - n_plain == 2 * n_gn + n_ln + n_res
    rel = ((fused.float() - plain.float()).norm() / plain.float().norm()).item()
    assert rel < 2e-2, rel
