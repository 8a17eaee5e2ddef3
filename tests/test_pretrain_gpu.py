"""Training side of hot path (ii) on the B200 (train/latent_wm_pretrain.py:164-217): gradients of the encoder, the noise layers
and the losses against autograd through the CPU oracle on identical inputs and parameters.  fp32 everywhere; tolerances are
elementwise (rtol 1e-4 / atol 1e-5 of the tensor's max unless stated)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, want, rtol=1e-4, atol_rel=1e-5, what=""):
    want = want.float().cpu()
    torch.testing.assert_close(got.float().cpu(), want, rtol=rtol, atol=atol_rel * float(want.abs().max()) + 1e-12, msg=lambda m: f"{what}: {m}")


def _layer_grads(layer_idx, params, x, gy, noise=None):
    from oracle import noise_oracle as NO

    xr = x.clone().requires_grad_(True)
    y = NO.apply_layer(xr, layer_idx, params, noise)
    y.backward(gy)
    return y.detach(), xr.grad


@pytest.mark.parametrize("shape", [(2, 3, 64, 80), (1, 3, 37, 50)])
def test_jpeg_backward(cuda_device, shape):
    from aqualora_b200 import noise_layers as NL

    g = torch.Generator().manual_seed(1)
    x = torch.rand(shape, generator=g) * 2 - 1
    gy = torch.randn(shape, generator=g)
    y_ref, gx_ref = _layer_grads(1, {}, x, gy)
    xc = x.to(cuda_device).requires_grad_(True)
    y = NL.jpeg_mask(xc)
    y.backward(gy.to(cuda_device))
    _close(y.detach(), y_ref, 1e-4, 2e-5, "jpeg y")
    _close(xc.grad, gx_ref, 1e-4, 2e-5, "jpeg gx")


def test_crop_resize_backward(cuda_device):
    from aqualora_b200 import noise_layers as NL

    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 96, 112, generator=g) * 2 - 1
    for p in (dict(top=5, left=9, crop_h=61, crop_w=70, resize_h=53, resize_w=88), dict(top=0, left=0, crop_h=96, crop_w=112, resize_h=120, resize_w=57)):
        gy = torch.randn(2, 3, 64, 72, generator=g)
        from oracle import noise_oracle as NO

        xr = x.clone().requires_grad_(True)
        NO.crop_resize(xr, out_hw=(64, 72), **p).backward(gy)
        xc = x.to(cuda_device).requires_grad_(True)
        NL.crop_resize(xc, out_hw=(64, 72), **p).backward(gy.to(cuda_device))
        _close(xc.grad, xr.grad, 1e-4, 1e-5, "crop_resize gx")           # atomics: order of the fp32 sums varies


def test_gauss_blur_backward(cuda_device):
    from aqualora_b200 import noise_layers as NL

    g = torch.Generator().manual_seed(3)
    for shape, sig, ks in (((2, 3, 40, 56), [0.7, 4.0], (3, 9)), ((1, 3, 9, 12), [2.5], (3, 5))):
        x = torch.rand(shape, generator=g) * 2 - 1
        gy = torch.randn(shape, generator=g)
        from oracle import noise_oracle as NO

        xr = x.clone().requires_grad_(True)
        NO.gaussian_blur(xr, sig, ks).backward(gy)
        xc = x.to(cuda_device).requires_grad_(True)
        NL.gaussian_blur(xc, sig, ks).backward(gy.to(cuda_device))
        _close(xc.grad, xr.grad, 1e-4, 1e-5, "blur gx")


def test_gauss_noise_backward_is_identity(cuda_device):
    from aqualora_b200 import noise_layers as NL

    x = torch.rand(2, 3, 16, 16, device=cuda_device).requires_grad_(True)
    gy = torch.randn(2, 3, 16, 16, device=cuda_device)
    NL.gaussian_noise(x, 0.1, 1234, 0).backward(gy)
    assert torch.equal(x.grad, gy)


def test_color_jiggle_backward(cuda_device):
    from aqualora_b200 import noise_layers as NL
    from oracle import noise_oracle as NO

    g = torch.Generator().manual_seed(4)
    rng = np.random.default_rng(4)
    for trial in range(4):
        x = torch.rand(2, 3, 24, 40, generator=g) * 2 - 1
        gy = torch.randn(2, 3, 24, 40, generator=g)
        p = NO.draw_params(rng, 5, 2)
        xr = x.clone().requires_grad_(True)
        NO.color_jiggle(xr, **p).backward(gy)
        xc = x.to(cuda_device).requires_grad_(True)
        y = NL.color_jiggle(xc, **p)
        y.backward(gy.to(cuda_device))
        # piecewise map: a pixel within rounding of a branch boundary (clamp edge, hue sector, channel tie) may take the other branch
        # in fp32 on the two machines; everywhere else the Jacobians agree to rounding
        diff = (xc.grad.cpu() - xr.grad).abs()
        tol = 1e-3 * xr.grad.abs().max() + 1e-3 * xr.grad.abs()
        assert (diff <= tol).float().mean().item() >= 0.999, (trial, p["order"], (diff > tol).sum().item())


@pytest.mark.parametrize("hw", [(64, 64), (96, 96), (40, 56)])
def test_secret_encoder_backward(cuda_device, hw):
    from aqualora_b200.models import SecretEncoder
    from oracle.pretrain_oracle import SecretEncoderRef

    torch.manual_seed(0)
    ref = SecretEncoderRef(48)
    with torch.no_grad():
        ref.secret_scaler[5].weight.normal_(0, 0.05)
        ref.secret_scaler[5].bias.normal_(0, 0.05)
    enc = SecretEncoder(48)
    enc.load_state_dict(ref.state_dict())
    enc = enc.to(cuda_device)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 4, *hw, generator=g)
    msg = torch.randint(0, 2, (3, 48), generator=g).float()
    g1, g2 = torch.randn(3, 4, *hw, generator=g), torch.randn(3, 4, *hw, generator=g)
    xr = x.clone().requires_grad_(True)
    xo_r, c_r = ref(xr, msg)
    (xo_r * g1).sum().add((c_r * g2).sum()).backward()
    xc = x.to(cuda_device).requires_grad_(True)
    xo, c = enc(xc, msg.to(cuda_device))
    ((xo * g1.to(cuda_device)).sum() + (c * g2.to(cuda_device)).sum()).backward()
    _close(c.detach(), c_r.detach(), 1e-5, 1e-6, "c")
    _close(xc.grad, xr.grad, 1e-5, 1e-6, "gx")
    for (n, p), (_, q) in zip(enc.named_parameters(), ref.named_parameters()):
        _close(p.grad, q.grad, 2e-4, 2e-5, n)
    # encode(): the map alone
    enc.zero_grad(); ref.zero_grad()
    cm = enc.encode(msg.to(cuda_device))
    cm.pow(2).sum().backward()
    ref.secret_scaler(msg).pow(2).sum().backward()
    for (n, p), (_, q) in zip(enc.named_parameters(), ref.named_parameters()):
        _close(p.grad, q.grad, 2e-4, 2e-5, "encode " + n)


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 3, 96, 80), (2, 3, 512, 512)])
def test_prvl_loss_and_gradient(cuda_device, shape):
    from aqualora_b200.losses import PRVL_loss
    from oracle.pretrain_oracle import prvl_loss

    g = torch.Generator().manual_seed(6)
    a = torch.rand(shape, generator=g) * 2 - 1
    b = a + 0.1 * torch.randn(shape, generator=g)
    br = b.clone().requires_grad_(True)
    want = prvl_loss(a, br)
    (want * 1.5).backward()
    bc = b.to(cuda_device).requires_grad_(True)
    got = PRVL_loss(a.to(cuda_device), bc)
    (got * 1.5).backward()
    assert abs(got.item() - want.item()) <= 2e-6 * abs(want.item()) + 1e-9, (got.item(), want.item())
    _close(bc.grad, br.grad, 1e-5, 1e-6, "prvl grad")
    assert int((bc.grad != 0).sum()) <= 3 * 32 * 32


def test_prvl_golden(cuda_device, golden_dir):
    """Values produced by the reference's own PRVL_loss source (tools/gen_golden.py)."""
    import os

    from aqualora_b200.losses import PRVL_loss

    gold = torch.load(os.path.join(golden_dir, "pretrain_small.pt"), weights_only=False)
    for c in gold["prvl"]:
        if c["a"] is None:
            continue
        got = PRVL_loss(c["a"].to(cuda_device), c["b"].to(cuda_device))
        assert abs(got.item() - c["value"].item()) <= 2e-6 * abs(c["value"].item()), (c["shape"], got.item(), c["value"].item())


def test_bce_with_logits(cuda_device):
    from aqualora_b200.losses import binary_cross_entropy_with_logits as bce

    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 48, 2, generator=g) * 3
    y = torch.nn.functional.one_hot(torch.randint(0, 2, (4, 48), generator=g), 2).float()
    xr = x.clone().requires_grad_(True)
    want = torch.nn.functional.binary_cross_entropy_with_logits(xr, y)
    want.backward()
    xc = x.to(cuda_device).requires_grad_(True)
    got = bce(xc, y.to(cuda_device))
    got.backward()
    assert abs(got.item() - want.item()) <= 1e-5 * want.item()
    _close(xc.grad, xr.grad, 1e-4, 1e-6, "bce grad")


def _assert_param_grads_close(got_mod, ref_mod, tol):
    """Per-tensor relative Frobenius error of every parameter gradient.  A convolution that feeds a batch-statistics BatchNorm has a
    gradient that is a difference of large cancelling terms (the loss is invariant to the scale and, per channel, the mean of its
    output), so tensors whose reference gradient is at the fp32 noise floor of that cancellation are compared absolutely against the
    typical gradient magnitude of the network instead."""
    gr = dict(ref_mod.named_parameters())
    rows = []
    for n, p in got_mod.named_parameters():
        q = gr[n].grad
        rows.append((n, float(q.norm()), float((p.grad.cpu() - q).norm()), q.numel()))
    typical = sorted(r[1] / r[3] ** 0.5 for r in rows)[len(rows) // 2]          # median per-element gradient magnitude
    bad = [(n, qn, dn) for n, qn, dn, k in rows if dn > tol * qn + tol * typical * k ** 0.5]
    assert not bad, (len(bad), sorted(bad, key=lambda r: -r[2] / (r[1] + 1e-30))[:6], typical)


def test_decoder_train_mode_matches_torchvision(cuda_device):
    """SecretDecoder.train(): batch-statistics BatchNorm forward + backward vs torchvision's efficientnet_b1 in train mode on the
    CPU (stochastic depth and dropout off on both sides: their masks come from different RNG streams)."""
    from aqualora_b200.decoder import SecretDecoder
    from oracle.pretrain_oracle import SecretDecoderRef

    torch.manual_seed(0)
    ref = SecretDecoderRef(48).train()
    for mod in ref.modules():
        if mod.__class__.__name__ == "StochasticDepth":
            mod.p = 0.0
    ref.model.classifier[0].p = 0.0
    dec = SecretDecoder(48)
    dec.load_state_dict(ref.state_dict())
    dec = dec.to(cuda_device).train()
    dec.stochastic_depth_prob, dec.dropout_p = 0.0, 0.0
    g = torch.Generator().manual_seed(8)
    x = torch.rand(2, 3, 512, 512, generator=g) * 2 - 1
    gy = torch.randn(2, 48, 2, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xc = x.to(cuda_device).requires_grad_(True)
    yc = dec(xc)
    yc.backward(gy.to(cuda_device))
    _close(yc.detach(), yr.detach(), 2e-3, 2e-3, "train logits")
    _close(xc.grad, xr.grad, 2e-2, 2e-2, "train gx")
    sd_r, sd_c = ref.state_dict(), dec.state_dict()
    for k in ("model.features.0.1.running_mean", "model.features.0.1.running_var", "model.features.8.1.running_var",
              "model.features.4.2.block.1.1.running_mean"):
        _close(sd_c[k], sd_r[k], 1e-3, 1e-3, k)
    _assert_param_grads_close(dec, ref, 5e-2)


def test_pretrain_step_matches_oracle(cuda_device):
    """configs[0] promoted to the GPU: one train/latent_wm_pretrain.py:164-217 iteration (B = 2, 48 bits, 64 x 64 latents) with the
    stub VAE / LPIPS of the oracle on both sides, every loss-schedule stage and three noise layers."""
    import random

    from aqualora_b200 import noise_layers as NL
    from aqualora_b200 import pretrain
    from aqualora_b200.decoder import SecretDecoder
    from aqualora_b200.models import SecretEncoder
    from oracle import pretrain_oracle as PO

    torch.manual_seed(1)
    enc_r = PO.SecretEncoderRef(48)
    with torch.no_grad():
        enc_r.secret_scaler[5].weight.normal_(0, 0.05)
    dec_r = PO.SecretDecoderRef(48).train()
    for mod in dec_r.modules():
        if mod.__class__.__name__ == "StochasticDepth":
            mod.p = 0.0
    dec_r.model.classifier[0].p = 0.0
    vae = PO.StubVAE(0)
    vae_c = PO.StubVAE(0).to(cuda_device)        # frozen third-party stand-in: plain torch on both sides
    g = torch.Generator().manual_seed(9)
    image = torch.rand(2, 3, 512, 512, generator=g) * 2 - 1
    msg = torch.randint(0, 2, (2, 48), generator=g)
    cases = [(0, {}, True, 0), (1, {}, False, 2), (3, {"sigmas": [1.5, 3.0]}, False, 1)]
    for layer, params, warmup, stage in cases:
        enc_c = SecretEncoder(48); enc_c.load_state_dict(enc_r.state_dict()); enc_c = enc_c.to(cuda_device).train()
        dec_c = SecretDecoder(48); dec_c.load_state_dict(dec_r.state_dict()); dec_c = dec_c.to(cuda_device).train()
        dec_c.stochastic_depth_prob, dec_c.dropout_p = 0.0, 0.0
        enc_r.zero_grad(); dec_r.zero_grad()
        out_r = PO.pretrain_step(enc_r, dec_r, vae, image, msg, layer, params, random.Random(3), warmup, stage)
        if layer == 0:
            override = lambda im: im
        elif layer == 1:
            override = NL.jpeg_mask
        else:
            override = lambda im: NL.gaussian_blur(im, params["sigmas"])
        out_c = pretrain.pretrain_step(enc_c, dec_c, vae_c.encode, vae_c.decode, PO.lpips_stub, None, image.to(cuda_device),
                                       msg.to(cuda_device), random.Random(3), None, warmup, stage, layer_override=override)
        for k in ("loss", "msgloss", "lpips", "prvl"):
            assert abs(out_c[k].item() - out_r[k].item()) <= 2e-3 * abs(out_r[k].item()) + 1e-7, (layer, k, out_c[k].item(), out_r[k].item())
        _close(out_c["wm_image"], out_r["wm_image"], 1e-4, 1e-5, "wm_image")
        for (n, p), (_, q) in zip(enc_c.named_parameters(), enc_r.named_parameters()):
            rel = ((p.grad.cpu() - q.grad).norm() / (q.grad.norm() + 1e-20)).item()
            assert rel < 5e-2, (layer, n, rel)
        _assert_param_grads_close(dec_c, dec_r, 5e-2)
        ck = pretrain.checkpoint_dict(enc_c, dec_c)
        assert set(ck) == {"sec_decoder", "sec_encoder"} and set(ck["sec_decoder"]) == set(dec_r.state_dict())


def test_stage3_decoder_finetune_step(cuda_device):
    """train/rob_enhance_finetune.py:1020-1038 (decoder only, train mode): distort -> decode -> BCE -> backward, vs the oracle."""
    import numpy as np

    from aqualora_b200 import noise_layers as NL
    from aqualora_b200 import pretrain
    from aqualora_b200.decoder import SecretDecoder
    from oracle import noise_oracle as NO
    from oracle.pretrain_oracle import SecretDecoderRef

    torch.manual_seed(2)
    ref = SecretDecoderRef(48).train()
    for mod in ref.modules():
        if mod.__class__.__name__ == "StochasticDepth":
            mod.p = 0.0
    ref.model.classifier[0].p = 0.0
    dec = SecretDecoder(48)
    dec.load_state_dict(ref.state_dict())
    dec = dec.to(cuda_device).train()
    dec.stochastic_depth_prob, dec.dropout_p = 0.0, 0.0
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 512, 512, generator=g)
    msg = torch.randint(0, 2, (2, 48), generator=g)
    # `blur` branch of distorsion_unit: sigma 4, kernel (3, 5) on [0, 1] images
    x_ref = (NO.gaussian_blur(img, [4.0, 4.0], (3, 5)) * 2 - 1).detach()
    logits = ref(x_ref)
    loss_ref = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.nn.functional.one_hot(msg, 2).float())
    loss_ref.backward()
    loss, acc = pretrain.decoder_finetune_step(dec, img.to(cuda_device), msg.to(cuda_device),
                                               lambda t: NL.distorsion_unit(t, "blur", rng=np.random.default_rng(0)))
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * loss_ref.item()
    assert 0.0 <= acc.item() <= 1.0
    _assert_param_grads_close(dec, ref, 5e-2)


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("C,k,s,H,W", [(16, 3, 1, 24, 24), (96, 3, 2, 33, 30), (144, 5, 2, 32, 32), (480, 5, 1, 16, 16), (32, 3, 1, 7, 9)])
def test_train_depthwise_kernels(cuda_device, C, k, s, H, W):
    """csrc/decoder_train.cu depthwise forward / input gradient / weight gradient vs the library convolution on the same device."""
    from aqualora_b200.decoder import _DwConvFn

    _no_tf32()
    g = torch.Generator(device=cuda_device).manual_seed(C + k)
    x = torch.randn(2, C, H, W, generator=g, device=cuda_device).contiguous(memory_format=torch.channels_last)
    w = torch.randn(C, 1, k, k, generator=g, device=cuda_device) * 0.2
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, None, s, (k - 1) // 2, 1, C)
    gy = torch.randn(yr.shape, generator=g, device=cuda_device).contiguous(memory_format=torch.channels_last)
    yr.backward(gy)
    xc, wc = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yc = _DwConvFn.apply(xc, wc, s)
    yc.backward(gy)
    _close(yc.detach(), yr.detach(), 1e-5, 1e-6, "dw y")
    _close(xc.grad, xr.grad, 1e-5, 1e-6, "dw gx")
    _close(wc.grad, wr.grad, 1e-4, 1e-5, "dw gw")


@pytest.mark.parametrize("K,N,H,with_se", [(16, 96, 16, False), (96, 24, 16, True), (320, 1280, 16, False), (1152, 192, 16, True), (32, 16, 32, True)])
def test_train_pointwise_kernels(cuda_device, K, N, H, with_se):
    from aqualora_b200.decoder import _PwConvFn

    _no_tf32()
    g = torch.Generator(device=cuda_device).manual_seed(K + N)
    B = 2
    x = torch.randn(B, K, H, H, generator=g, device=cuda_device).contiguous(memory_format=torch.channels_last)
    w = torch.randn(N, K, 1, 1, generator=g, device=cuda_device) * K ** -0.5
    se = torch.rand(B, K, 1, 1, generator=g, device=cuda_device) if with_se else None
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ser = se.clone().requires_grad_(True) if with_se else None
    yr = torch.nn.functional.conv2d(xr * ser if with_se else xr, wr)
    gy = torch.randn(yr.shape, generator=g, device=cuda_device).contiguous(memory_format=torch.channels_last)
    yr.backward(gy)
    xc, wc = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    sec = se.clone().requires_grad_(True) if with_se else None
    yc = _PwConvFn.apply(xc, wc, sec)
    yc.backward(gy)
    # 3 x TF32 == fp32 FFMA accuracy; the library sums the K <= 1152 products in another order (bound ~ sqrt(K) * 2^-24 * sum |x||w|)
    _close(yc.detach(), yr.detach(), 1e-4, 2e-5, "pw y")
    _close(xc.grad, xr.grad, 1e-4, 2e-5, "pw gx")
    _close(wc.grad, wr.grad, 1e-4, 1e-5, "pw gw")
    if with_se:
        _close(sec.grad, ser.grad, 1e-4, 1e-5, "pw g_se")


def test_train_stem_kernels(cuda_device):
    from aqualora_b200.decoder import _StemConvFn

    _no_tf32()
    g = torch.Generator(device=cuda_device).manual_seed(11)
    for (H, W) in ((64, 64), (37, 50)):
        x = torch.randn(2, 3, H, W, generator=g, device=cuda_device)
        w = torch.randn(32, 3, 3, 3, generator=g, device=cuda_device) * 0.2
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        yr = torch.nn.functional.conv2d(xr, wr, None, 2, 1)
        gy = torch.randn(yr.shape, generator=g, device=cuda_device).contiguous(memory_format=torch.channels_last)
        yr.backward(gy)
        xc, wc = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        yc = _StemConvFn.apply(xc, wc)
        yc.backward(gy)
        _close(yc.detach(), yr.detach(), 1e-5, 1e-6, "stem y")
        _close(xc.grad, xr.grad, 1e-5, 1e-6, "stem gx")
        _close(wc.grad, wr.grad, 1e-4, 1e-5, "stem gw")


@pytest.mark.parametrize("C,M_hw,act", [(16, 64 * 64, True), (96, 33 * 31, True), (1920, 16 * 16, False), (1152, 8 * 8, True)])
def test_train_batchnorm_kernels(cuda_device, C, M_hw, act):
    """csrc/bn_train.cu vs F.batch_norm(training=True) [+ SiLU] incl. the running-statistics update and all three gradients."""
    from aqualora_b200.decoder import _BnActFn

    g = torch.Generator(device=cuda_device).manual_seed(C)
    side = int(M_hw ** 0.5)
    H, W = side, M_hw // side
    z = (torch.randn(3, C, H, W, generator=g, device=cuda_device) * 2 + 0.5).contiguous(memory_format=torch.channels_last)
    gamma = torch.rand(C, generator=g, device=cuda_device) + 0.5
    beta = torch.randn(C, generator=g, device=cuda_device) * 0.1
    rm0, rv0 = torch.randn(C, generator=g, device=cuda_device) * 0.1, torch.rand(C, generator=g, device=cuda_device) + 0.5
    zr, gr, br = z.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = rm0.clone(), rv0.clone()
    yr = torch.nn.functional.batch_norm(zr, rm_r, rv_r, gr, br, True, 0.1, 1e-5)
    yr = torch.nn.functional.silu(yr) if act else yr
    gy = torch.randn(yr.shape, generator=g, device=cuda_device).contiguous(memory_format=torch.channels_last)
    yr.backward(gy)
    zc, gc, bc = z.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_c, rv_c = rm0.clone(), rv0.clone()
    yc = _BnActFn.apply(zc, gc, bc, rm_c, rv_c, 1e-5, 0.1, act)
    yc.backward(gy)
    _close(yc.detach(), yr.detach(), 1e-4, 1e-5, "bn y")
    _close(rm_c, rm_r, 1e-5, 1e-6, "running_mean")
    _close(rv_c, rv_r, 1e-5, 1e-6, "running_var")
    _close(zc.grad, zr.grad, 1e-3, 1e-4, "bn gz")
    _close(gc.grad, gr.grad, 1e-3, 1e-4, "bn g_gamma")
    _close(bc.grad, br.grad, 1e-3, 1e-4, "bn g_beta")
