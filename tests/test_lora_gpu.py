"""Hot path (i) on a B200: the fused projection + watermark-LoRA kernels, called through the drop-in modules and the
C ABI (ctypes -> libaqualora_b200.so), against

  * the golden vectors produced by the REFERENCE's own utils/lora_modules.py (tests/golden/lora_*.pt),
  * the CPU oracle (oracle/lora_oracle.py) on seeded inputs, including the reference's bf16-autocast arithmetic,
  * size-independent properties at the BASELINE shapes (zero scale == base bit-for-bit, linearity in the message scale,
    a plain PyTorch fp32 matmul of the same operands on the device).

Tolerance: activations are bf16 (8-bit mantissa, the precision the reference trains in: train/ppft_train.py:569-581).
A bf16-rounded output of an fp32-accumulated contraction differs from the fp32 result by <= 2^-8 relative per rounding;
the kernel rounds H, Hs and Y (the reference's autocast rounds the same three plus the base output and the LoRA output
separately), so |y - y_fp32| <= 2e-2 * max|y| is the stated bound for y / gx, and 2e-2 relative (Frobenius) for the fp32
weight gradients accumulated from bf16 operands.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _max_rel(got, want):
    want = want.float()
    return ((got.float().cpu() - want.cpu()).abs().max() / (want.abs().max() + 1e-12)).item()


def _elem_err(got, want, rtol=2e-2, atol_rms=4e-2):
    """Worst element of |got - want| / (rtol * |want| + atol_rms * rms(want)): < 1 means EVERY element is within 2 % of its own
    magnitude plus 4 % of the tensor's rms.  (bf16 operands against the fp32 golden: each of the K products carries ~2^-8 of
    relative rounding, so an output element is off by ~0.5 % of the output rms with random sign -- the absolute part is a > 5 sigma
    allowance for elements near zero, the relative part covers the output's own bf16 rounding and large elements.  A bound
    normalised by the tensor's MAXIMUM, as used before, let small elements be wrong by many times their value.)"""
    want = want.float().cpu()
    got = got.float().cpu()
    rms = want.pow(2).mean().sqrt()
    return ((got - want).abs() / (rtol * want.abs() + atol_rms * rms + 1e-30)).max().item()


def _fro_rel(got, want):
    want = want.float().cpu()
    return ((got.float().cpu() - want).norm() / (want.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def lm(cuda_device):
    from aqualora_b200 import _lib, lora_modules

    assert _lib.load().aq_arch() == 100
    return lora_modules


def _build_linear(lm, c, dev):
    dout, din = c["w"].shape
    lin = lm.LoRACompatibleLinear(din, dout, bias=c["b"] is not None)
    lin.weight.data.copy_(c["w"])
    if c["b"] is not None:
        lin.bias.data.copy_(c["b"])
    lin = lin.to(dev, torch.bfloat16)
    lin.requires_grad_(False)
    lora = lm.LoRALinearLayer(din, dout, c["r"], network_alpha=c["alpha"])
    lora.down.weight.data.copy_(c["down"])
    lora.up.weight.data.copy_(c["up"])
    lora = lora.to(dev)           # fp32 master weights, as in train/ppft_train.py:651-666
    lin.set_lora_layer(lora)
    return lin, lora


def test_linear_golden_forward_backward(lm, cuda_device, golden_dir):
    """Reference outputs/grads (fp32) vs the CUDA path on the bf16-rounded operands."""
    dev = cuda_device
    for c in _load(golden_dir, "lora_linear.pt"):
        lin, lora = _build_linear(lm, c, dev)
        x = c["x"].to(dev, torch.bfloat16).requires_grad_(True)
        scale = c["scale"]
        if isinstance(scale, torch.Tensor):
            scale = scale.to(dev).requires_grad_(True)
        y = lin(x, scale)
        assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(c["y"].shape)
        assert _elem_err(y, c["y"]) < 1.0, c["kind"]
        y.backward(c["gy"].to(dev, torch.bfloat16))
        assert _elem_err(x.grad, c["gx"]) < 1.0
        assert _fro_rel(lora.down.weight.grad, c["g_down"]) < BF16_TOL
        assert _fro_rel(lora.up.weight.grad, c["g_up"]) < BF16_TOL
        if isinstance(scale, torch.Tensor):
            assert _fro_rel(scale.grad, c["g_scale"]) < BF16_TOL
        # lora_layer is None -> exact base op (utils/lora_modules.py:57-59)
        lin.set_lora_layer(None)
        yb = lin(x.detach(), 1.0)
        assert _elem_err(yb, c["y_base"]) < 1.0


def test_conv1x1_golden_forward_backward(lm, cuda_device, golden_dir):
    dev = cuda_device
    for c in _load(golden_dir, "lora_conv1x1.pt"):
        co, ci = c["w"].shape[:2]
        conv = lm.LoRACompatibleConv(ci, co, kernel_size=1)
        conv.weight.data.copy_(c["w"]); conv.bias.data.copy_(c["b"])
        conv = conv.to(dev, torch.bfloat16)
        conv.requires_grad_(False)
        lora = lm.LoRAConv2dLayer(ci, co, rank=c["r"], network_alpha=c["alpha"])
        lora.down.weight.data.copy_(c["down"]); lora.up.weight.data.copy_(c["up"])
        lora = lora.to(dev)
        conv.set_lora_layer(lora)
        x = c["x"].to(dev, torch.bfloat16).requires_grad_(True)
        scale = c["scale"].to(dev).requires_grad_(True) if isinstance(c["scale"], torch.Tensor) else c["scale"]
        y = conv(x, scale)
        assert tuple(y.shape) == tuple(c["y"].shape)
        assert _elem_err(y, c["y"]) < 1.0
        y.backward(c["gy"].to(dev, torch.bfloat16))
        assert _elem_err(x.grad, c["gx"]) < 1.0
        assert _fro_rel(lora.down.weight.grad, c["g_down"]) < BF16_TOL
        assert _fro_rel(lora.up.weight.grad, c["g_up"]) < BF16_TOL
        if isinstance(scale, torch.Tensor):
            assert _fro_rel(scale.grad, c["g_scale"]) < BF16_TOL


@pytest.mark.parametrize("r,B,N,din,dout,bias", [(320, 2, 300, 320, 320, True), (128, 3, 77, 768, 640, False), (72, 2, 256, 640, 1280, True),
                                                  (4, 2, 64, 64, 96, True), (12, 1, 40, 128, 64, False)])
def test_other_ranks_forward_backward(lm, cuda_device, r, B, N, din, dout, bias):
    """Ranks beyond one 64-wide slice (the reference's released recipe is rank 320, train/README.md:34-48: chunked by the ABI layer),
    ragged ranks, and ranks that are not a multiple of 8 (ppft_train.py's default --rank 4: zero-padded), against the closed form on
    the bf16-rounded operands."""
    from oracle import lora_oracle as O

    dev = cuda_device
    g = torch.Generator().manual_seed(100 + r)
    r16 = lambda t: t.bfloat16().float()
    x = r16(torch.randn(B, N, din, generator=g))
    w = r16(torch.randn(dout, din, generator=g) * din ** -0.5)
    b = r16(torch.randn(dout, generator=g) * 0.1) if bias else None
    dn = r16(torch.randn(r, din, generator=g) * din ** -0.5)
    up = r16(torch.randn(dout, r, generator=g) * r ** -0.5)
    sc = r16(1 + 0.5 * torch.randn(B, r, generator=g))
    gy = r16(torch.randn(B, N, dout, generator=g) * 0.1)
    want = O.closed_form_linear(x, w, b, dn, up, sc)
    dx, dd, du, ds = O.closed_form_linear_grads(x, w, dn, up, sc, gy)
    lin = lm.LoRACompatibleLinear(din, dout, bias=bias)
    lin.weight.data.copy_(w)
    if bias:
        lin.bias.data.copy_(b)
    lin = lin.to(dev, torch.bfloat16)
    lin.requires_grad_(False)
    lora = lm.LoRALinearLayer(din, dout, r)
    lora.down.weight.data.copy_(dn); lora.up.weight.data.copy_(up)
    lin.set_lora_layer(lora.to(dev))
    xd = x.to(dev, torch.bfloat16).requires_grad_(True)
    sd = sc.to(dev).requires_grad_(True)
    y = lin(xd, sd)
    y.backward(gy.to(dev, torch.bfloat16))
    assert _elem_err(y, want) < 1.0
    assert _elem_err(xd.grad, dx) < 1.0
    assert _fro_rel(lora.down.weight.grad, dd) < BF16_TOL and _fro_rel(lora.up.weight.grad, du) < BF16_TOL
    assert _fro_rel(sd.grad, ds) < BF16_TOL
    assert tuple(lora.down.weight.grad.shape) == (r, din) and tuple(sd.grad.shape) == (B, r)
    # zero diagonal -> bit-identical to the base op also through the chunked launches
    y0 = lin(xd.detach(), torch.zeros(B, r, device=dev))
    lin.set_lora_layer(None)
    assert torch.equal(y0, lin(xd.detach(), 1.0))


@pytest.mark.parametrize("M,K,N,r,tok", [(4096, 320, 320, 64, 1024), (1000, 328, 200, 8, 500), (2048, 1280, 640, 320, 1024)])
def test_residual_rides_in_the_epilogue(lm, cuda_device, M, K, N, r, tok):
    """aq_lora_linear_fwd_residual == projection, then `+ residual` (one bf16 rounding of the sum instead of two), and the module-level
    helper routes the residual's gradient through unchanged."""
    from aqualora_b200 import ops

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(M + N)
    x = torch.randn(M, K, generator=g, device=dev).bfloat16()
    w = (torch.randn(N, K, generator=g, device=dev) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g, device=dev).bfloat16()
    dn = (torch.randn(r, K, generator=g, device=dev) * K ** -0.5).bfloat16()
    up = (torch.randn(N, r, generator=g, device=dev) * 0.1).bfloat16()
    sc = torch.randn(M // tok, r, generator=g, device=dev)
    res = torch.randn(M, N, generator=g, device=dev).bfloat16()
    y_plain, _ = ops.lora_linear_fwd(x, w, b, dn, up, sc, tok)
    y_res, _ = ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, residual=res)
    want = y_plain.float() + res.float()
    # bf16 roundings: one per rank chunk on either side (ranks above 64 accumulate Y in place, chunk by chunk) + the unfused add
    # Ranks above 64 accumulate Y in place chunk by chunk, each pass rounding to bf16: the two sides then differ by a few ulps of the
    # LARGEST intermediate sum, not of the final value -- compare against the fp32 closed form instead, like the other rank tests.
    if r > 64:
        want = (x.float() @ w.float().t() + b.float() + ((x.float() @ dn.float().t()).view(M // tok, tok, r) * sc[:, None, :]).view(M, r) @ up.float().t()
                + res.float())
        assert _elem_err(y_res, want) < 1.0
    else:
        assert ((y_res.float() - want).abs() <= 2 ** -7 * (y_plain.float().abs() + want.abs()) + 1e-6).all()
    y0, _ = ops.lora_linear_fwd(x, w, None, None, None, None, tok, residual=res)          # plain projection + residual
    p0, _ = ops.lora_linear_fwd(x, w, None, None, None, None, tok)
    assert ((y0.float() - (p0.float() + res.float())).abs() <= 2 ** -7 * (p0.float().abs() + (p0.float() + res.float()).abs()) + 1e-6).all()
    if r <= 64:
        lin = lm.LoRACompatibleLinear(K, N).to(dev, torch.bfloat16)
        lin.weight.data.copy_(w); lin.bias.data.copy_(b)
        lin.requires_grad_(False)
        lora = lm.LoRALinearLayer(K, N, r)
        lora.down.weight.data.copy_(dn.float()); lora.up.weight.data.copy_(up.float())
        lin.set_lora_layer(lora.to(dev))
        B = M // tok
        xin = x.view(B, tok, K).clone().requires_grad_(True)
        rin = res.view(B, tok, N).clone().requires_grad_(True)
        y = lm.linear_with_residual(lin, xin, sc, rin)
        gy = torch.randn(B, tok, N, generator=g, device=dev).bfloat16()
        y.backward(gy)
        assert torch.equal(rin.grad, gy)
        xin2 = x.view(B, tok, K).clone().requires_grad_(True)
        (lin(xin2, sc) + rin.detach()).backward(gy)
        assert torch.equal(xin.grad, xin2.grad)


def test_fp16_and_fp32_activations_cast_at_the_boundary(lm, cuda_device):
    """The reference's README recipe runs fp16 (train/README.md:34-48): fp16 / fp32 activations and base weights are rounded to bf16 at
    the boundary and the result is returned in the caller's dtype; agreement is to bf16 rounding (stated in lora_modules._check_input)."""
    from oracle import lora_oracle as O

    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    B, N, din, dout, r = 2, 96, 320, 640, 64
    x = torch.randn(B, N, din, generator=g)
    w = torch.randn(dout, din, generator=g) * din ** -0.5
    b = torch.randn(dout, generator=g) * 0.1
    dn = torch.randn(r, din, generator=g) * din ** -0.5
    up = torch.randn(dout, r, generator=g) * 0.1
    sc = 1 + 0.5 * torch.randn(B, r, generator=g)
    want = O.closed_form_linear(x, w, b, dn, up, sc)
    for dt in (torch.float16, torch.float32):
        lin = lm.LoRACompatibleLinear(din, dout)
        lin.weight.data.copy_(w); lin.bias.data.copy_(b)
        lin = lin.to(dev, dt)
        lin.requires_grad_(False)
        lora = lm.LoRALinearLayer(din, dout, r)
        lora.down.weight.data.copy_(dn); lora.up.weight.data.copy_(up)
        lin.set_lora_layer(lora.to(dev))
        xd = x.to(dev, dt).requires_grad_(True)
        y = lin(xd, sc.to(dev))
        assert y.dtype == dt
        assert _elem_err(y, want) < 1.0
        y.float().pow(2).sum().backward()
        assert xd.grad is not None and xd.grad.dtype == dt and torch.isfinite(xd.grad).all()
        assert lora.down.weight.grad is not None and torch.isfinite(lora.down.weight.grad).all()


def test_standalone_lora_layer_matches_oracle(lm, cuda_device):
    """CustomLoRALinearLayerforward called on its own (utils/lora_modules.py:9-26), float and tensor scale."""
    from oracle import lora_oracle as O

    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 40, 64, generator=g).bfloat16()
    lora = lm.LoRALinearLayer(64, 96, 16, network_alpha=8.0)
    lora.up.weight.data.copy_(torch.randn(96, 16, generator=g) * 0.1)
    for scale in (0.5, torch.rand(2, 16, generator=g) + 0.5):
        want = O.lora_linear_layer_forward(x.float(), lora.down.weight.data.bfloat16().float(), lora.up.weight.data.bfloat16().float(),
                                           scale.bfloat16().float() if isinstance(scale, torch.Tensor) else scale, 8.0, 16)
        got = lora.to(dev)(x.to(dev), scale.to(dev) if isinstance(scale, torch.Tensor) else scale)
        assert _elem_err(got, want) < 1.0
        lora = lora.cpu()


@pytest.mark.parametrize("M,K,N,r,tok", [
    (128, 64, 64, 64, 128),          # one tile
    (1000, 328, 200, 8, 500),        # ragged everything
    (1232, 768, 320, 64, 77),        # cross-attention K/V: 77 tokens per sample, samples straddle row tiles
    (64 * 16, 1280, 1280, 64, 64),   # mid block: 64 tokens per sample, 2 samples per 128-row tile
    (4096, 320, 2560, 64, 1024),     # GEGLU projection, many column tiles per row block
    (2048, 2560, 640, 32, 1024),     # long K
])
def test_forward_matches_autocast_oracle(cuda_device, M, K, N, r, tok):
    """Against the reference's bf16-autocast arithmetic restated on CPU (oracle.bf16_autocast_linear)."""
    from aqualora_b200 import ops
    from oracle import lora_oracle as O

    dev = cuda_device
    g = torch.Generator().manual_seed(M + K + N)
    B = M // tok
    x = torch.randn(B, tok, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g).bfloat16()
    dn = torch.randn(r, K, generator=g) * K ** -0.5
    up = torch.randn(N, r, generator=g) * 0.1
    sc = (1 + 0.7 * torch.randn(B, r, generator=g)).bfloat16()
    want, h_want = O.bf16_autocast_linear(x, w, b, dn, up, sc)
    y, h = ops.lora_linear_fwd(x.reshape(M, K).to(dev), w.to(dev), b.to(dev), dn.bfloat16().to(dev), up.bfloat16().to(dev),
                               sc.float().to(dev), tok, save_h=True)
    assert _elem_err(y, want.reshape(M, N)) < 1.0
    # H is a single fp32-accumulated contraction rounded once: at most 1 bf16 ulp from the oracle's rounding
    assert _max_rel(h, h_want.reshape(M, r)) < 2 ** -7


def test_zero_scale_is_bit_identical_to_base(cuda_device):
    """scale = 0 -> the LoRA branch contributes exactly 0 (train/ppft_train.py:1026-1029 relies on it)."""
    from aqualora_b200 import ops

    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    for (M, K, N, r, tok) in [(4096, 320, 320, 64, 1024), (16 * 4096, 320, 320, 64, 4096), (1232, 768, 640, 64, 77)]:
        x = torch.randn(M, K, generator=g).bfloat16().to(dev)
        w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(dev)
        b = torch.randn(N, generator=g).bfloat16().to(dev)
        dn = (torch.randn(r, K, generator=g)).bfloat16().to(dev)
        up = (torch.randn(N, r, generator=g)).bfloat16().to(dev)
        zero = torch.zeros(M // tok, r, device=dev)
        y0, _ = ops.lora_linear_fwd(x, w, b, dn, up, zero, tok)
        yb, _ = ops.lora_linear_fwd(x, w, b, None, None, None, tok)
        assert torch.equal(y0, yb)


def test_full_size_properties(cuda_device):
    """BASELINE shape (B=16, 4096 tokens, 320 -> 320, r = 64): device fp32 matmul reference + linearity in the scale."""
    from aqualora_b200 import ops

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(2)
    B, tok, K, N, r = 16, 4096, 320, 320, 64
    M = B * tok
    x = torch.randn(M, K, generator=g, device=dev).bfloat16()
    w = (torch.randn(N, K, generator=g, device=dev) * K ** -0.5).bfloat16()
    dn = (torch.randn(r, K, generator=g, device=dev) * K ** -0.5).bfloat16()
    up = (torch.randn(N, r, generator=g, device=dev) * 0.1).bfloat16()
    s = (1 + 0.5 * torch.randn(B, r, generator=g, device=dev)).bfloat16().float()
    y, h = ops.lora_linear_fwd(x, w, None, dn, up, s, tok, save_h=True)
    h_ref = (x.float() @ dn.float().t()).bfloat16().float()
    hs = (h_ref * s.repeat_interleave(tok, 0)).bfloat16().float()
    ref = x.float() @ w.float().t() + hs @ up.float().t()
    assert ((y.float() - ref).abs().max() / ref.abs().max()).item() < BF16_TOL
    # linearity: y(2s) - y(0) == 2 (y(s) - y(0)) up to the bf16 roundings of Hs and Y
    y2, _ = ops.lora_linear_fwd(x, w, None, dn, up, 2 * s, tok)
    y0, _ = ops.lora_linear_fwd(x, w, None, dn, up, torch.zeros_like(s), tok)
    lhs = (y2.float() - y0.float())
    rhs = 2 * (y.float() - y0.float())
    assert ((lhs - rhs).abs().max() / ref.abs().max()).item() < BF16_TOL
    # every column tile width / grouping computes the same thing
    from aqualora_b200 import ops as _ops
    try:
        for bn, grp in [(64, 1), (128, 2), (160, 1), (192, 2)]:
            _ops.set_tuning(bn, grp)
            yt, _ = _ops.lora_linear_fwd(x, w, None, dn, up, s, tok)
            assert ((yt.float() - y.float()).abs().max() / ref.abs().max()).item() < 2 ** -7, (bn, grp)
    finally:
        _ops.set_tuning(0, 0)


@pytest.mark.parametrize("M,K,N,r,tok,dx", [
    (1024, 320, 320, 64, 256, True),
    (1232, 768, 320, 64, 77, False),     # text-context input needs no gradient (attn2.to_k / to_v)
    (2048, 640, 2560, 32, 1024, True),
    (1000, 328, 200, 8, 250, True),
])
def test_backward_matches_closed_form_oracle(cuda_device, M, K, N, r, tok, dx):
    from aqualora_b200 import ops
    from oracle import lora_oracle as O

    dev = cuda_device
    g = torch.Generator().manual_seed(7 * M + r)
    B = M // tok
    x = torch.randn(B, tok, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    dn = (torch.randn(r, K, generator=g) * K ** -0.5).bfloat16()
    up = (torch.randn(N, r, generator=g) * 0.1).bfloat16()
    sc = (1 + 0.7 * torch.randn(B, r, generator=g)).bfloat16().float()
    gy = (torch.randn(B, tok, N, generator=g) * 0.05).bfloat16()
    dxr, ddr, dur, dsr = O.closed_form_linear_grads(x.float(), w.float(), dn.float(), up.float(), sc, gy.float())
    xd, wd, dnd, upd, scd, gyd = (t.to(dev) for t in (x.reshape(M, K), w, dn, up, sc, gy.reshape(M, N)))
    _, h = ops.lora_linear_fwd(xd, wd, None, dnd, upd, scd, tok, save_h=True)
    g_dn = torch.zeros(r, K, device=dev); g_up = torch.zeros(N, r, device=dev); g_sc = torch.zeros(B, r, device=dev)
    gx = ops.lora_linear_bwd(gyd, xd, wd.t().contiguous() if dx else None, dnd.t().contiguous(), upd.t().contiguous(), scd, h,
                             g_dn, g_up, g_sc, tok)
    if dx:
        assert _elem_err(gx, dxr.reshape(M, K)) < 1.0
    else:
        assert gx is None
    assert _fro_rel(g_dn, ddr) < BF16_TOL and _fro_rel(g_up, dur) < BF16_TOL and _fro_rel(g_sc, dsr) < BF16_TOL
    # gradients ACCUMULATE into the caller's buffers (flat gradient buffer contract)
    ops.lora_linear_bwd(gyd, xd, wd.t().contiguous() if dx else None, dnd.t().contiguous(), upd.t().contiguous(), scd, h,
                        g_dn, g_up, g_sc, tok)
    assert _fro_rel(g_up, 2 * dur) < BF16_TOL


def test_wgrad_contraction(cuda_device):
    from aqualora_b200 import ops

    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    for (M, I, J, t) in [(1024, 320, 64, False), (5000, 1280, 64, True), (1232, 768, 16, True), (65536, 320, 64, False)]:
        p = torch.randn(M, I, generator=g).bfloat16()
        q = torch.randn(M, J, generator=g).bfloat16()
        c = torch.zeros((J, I) if t else (I, J), device=dev)
        ops.wgrad_tn(p.to(dev), q.to(dev), c, transpose_out=t)
        ref = p.double().t() @ q.double()
        assert _fro_rel(c, ref.t() if t else ref) < 1e-5      # fp32 accumulation of exact bf16 products


def test_mapper_and_flat_adamw(cuda_device, golden_dir):
    from aqualora_b200 import ops

    dev = cuda_device
    gm = _load(golden_dir, "models_small.pt")["mapper"]
    got = ops.mapper_fwd(gm["msg"].to(dev), gm["emb"].to(dev), round_bf16=False)
    torch.testing.assert_close(got.cpu(), gm["scale"], rtol=1e-6, atol=1e-6)
    # backward: dE = msg^T ds / sqrt(bits)
    ds = torch.randn(gm["msg"].shape[0], gm["emb"].shape[1], generator=torch.Generator().manual_seed(0))
    ge = torch.zeros_like(gm["emb"]).to(dev)
    ops.mapper_bwd(gm["msg"].to(dev), ds.to(dev), ge)
    torch.testing.assert_close(ge.cpu(), gm["msg"].t() @ ds / gm["emb"].shape[0] ** 0.5, rtol=1e-5, atol=1e-6)

    # clip_grad_norm_(1.0) + AdamW + zero_grad (train/ppft_train.py:1065-1068) vs torch on CPU
    n = 100_003
    g = torch.Generator().manual_seed(4)
    p0 = torch.randn(n, generator=g); gr = torch.randn(n, generator=g) * 0.05
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    pd, gd = p0.to(dev), torch.zeros(n, device=dev)
    m, v, nsq = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(1, device=dev)
    for step in (1, 2, 3):
        p_ref.grad = gr.clone() * step
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        gd.copy_(gr * step); nsq.zero_()
        ops.flat_sumsq(gd, nsq)
        ops.flat_clip_adamw(pd, gd, m, v, nsq, grad_scale=1.0, max_norm=1.0, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8,
                            weight_decay=1e-2, step=step)
        assert bool((gd == 0).all())
    torch.testing.assert_close(pd.cpu(), p_ref.detach(), rtol=1e-5, atol=1e-6)


def test_tiny_unet_ppft_step_matches_oracle(cuda_device):
    """One PPFT step on the tiny U-Net: CUDA path (bf16) vs the reference op sequence on CPU in fp32 (oracle patch)."""
    from aqualora_b200 import lora_modules, ppft
    from aqualora_b200.unet import UNetConfig, lora_target_keys
    from oracle import lora_oracle as O
    from oracle.patch import patch_with_oracle

    dev = cuda_device
    cfg = UNetConfig.tiny(16)
    rank, bits, B = 8, 48, 2
    unet = ppft.build_unet(cfg, dev, seed=3)
    emb = O.mapper_init(bits, rank, generator=torch.Generator().manual_seed(5))
    tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=rank, msg_bits=bits), emb, dev, lora_up_std=0.05, seed=1)
    g = torch.Generator().manual_seed(11)
    s = cfg.sample_size
    lat = torch.randn(B, 4, s, s, generator=g) * 0.18215
    wm = torch.randn(B, 4, s, s, generator=g) * 0.02 * 0.18215
    noise = torch.randn(B, 4, s, s, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g)
    msg = torch.randint(0, 2, (B, bits), generator=g).float()
    bf = lambda x: x.to(dev, torch.bfloat16)
    loss = tr.forward_backward(bf(lat), bf(wm), bf(noise), t.to(dev), bf(ctx), msg.to(dev))
    g_flat = tr.state.grad.clone()

    cpu_unet = ppft.build_unet(cfg, "cpu", dtype=torch.float32, seed=3)
    layers = lora_modules.inject_lora(cpu_unet, lora_target_keys(cpu_unet), rank)
    cpu_unet.load_state_dict({k: v.detach().float().cpu() for k, v in unet.state_dict().items()})
    patch_with_oracle(cpu_unet)
    for _, _, l in layers:
        l.down.weight.requires_grad_(True); l.up.weight.requires_grad_(True)
    E = tr.state.mapper_emb.detach().cpu().clone().requires_grad_(True)
    r16 = lambda x: x.bfloat16().float()
    scale = r16(O.mapper_forward(msg, E))
    ac = ppft.scaled_linear_alphas_cumprod()
    noisy = ppft.add_noise(ac, r16(lat), r16(noise), t)
    noisy_wm = ppft.add_noise(ac, r16(lat) + r16(wm), r16(noise), t)
    with torch.no_grad():
        clean = cpu_unet(noisy, t, r16(ctx), cross_attention_kwargs={"scale": torch.zeros_like(scale)}).sample
    pred = cpu_unet(noisy_wm, t, r16(ctx), cross_attention_kwargs={"scale": scale}).sample
    loss_ref = torch.nn.functional.mse_loss(pred, clean)
    loss_ref.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for _, _, l in layers for p in (l.down.weight, l.up.weight)])
    g_got = g_flat[:g_ref.numel()].cpu()
    cos = torch.nn.functional.cosine_similarity(g_got, g_ref, dim=0).item()
    # whole-network bf16 (activations, attention, norms) vs fp32: direction must agree, magnitude within 10 %
    assert cos > 0.995, cos
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 0.05
    assert abs(g_got.norm().item() / g_ref.norm().item() - 1) < 0.1
    ge = tr.state.mapper_grad.cpu().reshape(-1)
    assert torch.nn.functional.cosine_similarity(ge, E.grad.reshape(-1), dim=0).item() > 0.995
    tr.optimizer_step()
    assert bool((tr.state.grad == 0).all())


def test_secret_encoder_golden(cuda_device, golden_dir):
    """SecretEncoder (utils/models.py:51-81) vs the reference's own outputs."""
    from aqualora_b200.models import SecretEncoder

    g = _load(golden_dir, "models_small.pt")
    enc = SecretEncoder(48)
    enc.load_state_dict(g["encoder_state"])
    enc = enc.to(cuda_device)
    for key in ("encoder_64x64", "encoder_96x96", "encoder_40x56"):
        c = g[key]
        xo, cm = enc(c["x"].to(cuda_device), c["msg"].to(cuda_device))
        torch.testing.assert_close(cm.cpu(), c["c"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(xo.cpu(), c["x_out"], rtol=1e-5, atol=1e-6)


def test_state_dict_keys_follow_reference(cuda_device):
    """pytorch_lora_weights.safetensors key naming (train/ppft_train.py:443-471)."""
    from aqualora_b200 import lora_modules, ppft
    from aqualora_b200.unet import UNetConfig, lora_target_keys

    unet = ppft.build_unet(UNetConfig.tiny(16), cuda_device, seed=0)
    keys = lora_target_keys(unet)
    lora_modules.inject_lora(unet, keys, 8)
    sd = lora_modules.unet_attn_processors_state_dict(unet, keys)
    assert len(sd) == 2 * 192
    assert "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.to_q_lora.down.weight" in sd
    assert "down_blocks.0.attentions.0.proj_in.lora.up.weight" in sd
    assert "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.lora.down.weight" in sd


def _projection_family(lm, dev, din, douts, r, seed, bias=False, with_lora=True):
    g = torch.Generator().manual_seed(seed)
    mods = []
    for dout in douts:
        lin = lm.LoRACompatibleLinear(din, dout, bias=bias)
        lin.weight.data.copy_(torch.randn(dout, din, generator=g) * din ** -0.5)
        if bias:
            lin.bias.data.copy_(torch.randn(dout, generator=g) * 0.1)
        lin = lin.to(dev, torch.bfloat16).requires_grad_(False)
        if with_lora:
            lora = lm.LoRALinearLayer(din, dout, r)
            lora.down.weight.data.copy_(torch.randn(r, din, generator=g) * din ** -0.5)
            lora.up.weight.data.copy_(torch.randn(dout, r, generator=g) * 0.1)
            lin.set_lora_layer(lora.to(dev))
        mods.append(lin)
    return mods


@pytest.mark.parametrize("B,tok,din,douts,r,bias,need_dx", [
    (2, 300, 320, (320, 320, 320), 64, False, True),          # q / k / v of a self-attention, ragged last row block
    (2, 77, 768, (320, 320, 640, 640, 1280, 1280), 64, False, False),   # hoisted cross-attention K / V on the text context
    (3, 128, 64, (64, 200), 16, True, True),                  # small rank, bias, partial column tile
])
def test_grouped_projections_equal_separate_launches(lm, cuda_device, B, tok, din, douts, r, bias, need_dx):
    """project_group == the same modules called one by one (each utils/lora_modules.py:56-62): forward bit-for-bit (the
    accumulation order over k-blocks does not depend on the column tile), gradients to bf16 accumulation-order noise, and
    both within the bf16 bound of the fp32 closed form."""
    from oracle import lora_oracle as O

    dev = cuda_device
    mods = _projection_family(lm, dev, din, douts, r, seed=21, bias=bias)
    g = torch.Generator().manual_seed(22)
    x0 = torch.randn(B, tok, din, generator=g).bfloat16()
    s0 = (1 + 0.7 * torch.randn(B, r, generator=g)).bfloat16().float()   # fp32 leaf holding bf16 values: its gradient stays fp32
    gys = [(torch.randn(B, tok, d, generator=g) * 0.05).bfloat16().to(dev) for d in douts]

    def run(grouped):
        x = x0.to(dev).requires_grad_(need_dx)
        s = s0.to(dev).requires_grad_(True)
        for m in mods:
            m.lora_layer.zero_grad(set_to_none=True)
        ys = lm.project_group(mods, x, s) if grouped else [m(x, s) for m in mods]
        torch.autograd.backward(ys, gys)
        grads = [p.grad.clone() for m in mods for p in (m.lora_layer.down.weight, m.lora_layer.up.weight)]
        return [y.detach() for y in ys], (x.grad.clone() if need_dx else None), s.grad.clone(), grads

    run(False)                                               # fills the operand caches (bf16 copies, W^T): not counted below
    n0 = _launches()
    ys_g, gx_g, gs_g, gr_g = run(True)
    n_grouped = _launches() - n0
    n0 = _launches()
    ys_s, gx_s, gs_s, gr_s = run(False)
    n_separate = _launches() - n0
    assert n_separate - n_grouped == len(douts) - 1          # one forward launch instead of len(douts)
    for yg, ys, m in zip(ys_g, ys_s, mods):
        assert torch.equal(yg, ys)
        want = O.closed_form_linear(x0.float(), m.weight.float().cpu(), None if m.bias is None else m.bias.float().cpu(),
                                    m.lora_layer.down.weight.detach().bfloat16().float().cpu(),
                                    m.lora_layer.up.weight.detach().bfloat16().float().cpu(), s0.float())
        assert _elem_err(yg, want) < 1.0
    if need_dx:
        assert _elem_err(gx_g, gx_s) < 1.0
    # d(scale) flows back through the reference's `.to(weight_dtype)` cast (utils/lora_modules.py:15-17): the grouped call rounds
    # the SUM over its projections to bf16 once, separate calls round each projection's term -> up to one bf16 ulp (2^-8) apart
    assert _fro_rel(gs_g, gs_s) < 2.0 ** -8
    for a, b in zip(gr_g, gr_s):
        assert _fro_rel(a, b) < 2e-3


def _launches():
    from aqualora_b200 import _lib

    return _lib.load().aq_launch_count()


def test_grouped_plain_and_fallback(lm, cuda_device):
    """lora_disabled() groups the plain base projections; a family that cannot share a launch (LoRA on some members only)
    falls back to one fused launch per module with identical results."""
    dev = cuda_device
    mods = _projection_family(lm, dev, 128, (128, 256), 16, seed=5, bias=True)
    x = torch.randn(2, 50, 128, generator=torch.Generator().manual_seed(6)).bfloat16().to(dev)
    s = torch.ones(2, 16, device=dev)
    with torch.no_grad():
        [m(x, s) for m in mods]                              # operand caches
    with lm.lora_disabled(), torch.no_grad():
        n0 = _launches()
        got = lm.project_group(mods, x, s)
        assert _launches() - n0 == 1
        want = [m(x, s) for m in mods]
    for a, b, m in zip(got, want, mods):
        assert torch.equal(a, b)
        ref = torch.nn.functional.linear(x.float(), m.weight.float(), m.bias.float())
        assert _elem_err(a, ref) < 1.0
    mods[1].set_lora_layer(None)
    with torch.no_grad():
        n0 = _launches()
        got = lm.project_group(mods, x, s)
        assert _launches() - n0 == 2
        want = [m(x, s) for m in mods]
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_operand_cache_ignores_recycled_ids(lm, cuda_device):
    """The bf16 operand cache is keyed by id(parameter); CPython hands the id of a collected tensor to the next one.  An entry
    left by another tensor (same id, same version / address / shape by coincidence) must not be served."""
    import weakref

    p = torch.randn(8, 64, device=cuda_device)
    other = torch.randn(8, 64, device=cuda_device)
    d, dt = lm._packed(other, 8, 64)
    lm._PACK_CACHE[id(p)] = ((p._version, p.data_ptr(), lm._PACK_EPOCH), d, dt, weakref.ref(other))
    d2, dt2 = lm._packed(p, 8, 64)
    assert torch.equal(d2.float(), p.bfloat16().float()) and torch.equal(dt2, d2.t())
    w = torch.randn(16, 64, device=cuda_device).bfloat16()
    lm._WT_CACHE[id(w)] = ((w._version, w.data_ptr()), dt, weakref.ref(other))
    assert torch.equal(lm._weight_t(w, 16, 64), w.t())


def test_cfg_doubled_sampling_forward_matches_oracle(cuda_device):
    """the sampling call of train/rob_enhance_finetune.py:995-1012 -- classifier-free guidance doubles the batch and the
    message diagonal rides along as `torch.cat([mapper(msg)] * 2) * 1.03`, on a non-square latent grid (the script draws
    height / width from 512 ... 768) -- through the tiny U-Net under no_grad, vs the reference op sequence on CPU in fp32."""
    from aqualora_b200 import lora_modules, ppft
    from aqualora_b200.unet import UNetConfig, lora_target_keys
    from oracle import lora_oracle as O
    from oracle.patch import patch_with_oracle

    dev = cuda_device
    cfg = UNetConfig.tiny(16)
    rank, bits, B = 8, 48, 2
    unet = ppft.build_unet(cfg, dev, seed=3)
    emb = O.mapper_init(bits, rank, generator=torch.Generator().manual_seed(5))
    tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=rank, msg_bits=bits), emb, dev, lora_up_std=0.05, seed=1)
    g = torch.Generator().manual_seed(21)
    lat = torch.randn(B, 4, 24, 16, generator=g)
    ctx = torch.randn(2 * B, 77, cfg.cross_attention_dim, generator=g)          # [uncond; cond]
    msg = torch.randint(0, 2, (B, bits), generator=g).float()
    t = torch.tensor([481], dtype=torch.long)
    r16 = lambda x: x.bfloat16().float()
    s = r16(O.mapper_forward(msg, tr.state.mapper_emb.detach().cpu()))
    s2 = r16(torch.cat([s] * 2) * 1.03)
    x2 = torch.cat([lat] * 2)
    with torch.no_grad():
        got = unet(x2.to(dev, torch.bfloat16), t.to(dev), ctx.to(dev, torch.bfloat16),
                   cross_attention_kwargs={"scale": s2.to(dev)}).sample
    assert got.shape == (2 * B, 4, 24, 16)

    cpu_unet = ppft.build_unet(cfg, "cpu", dtype=torch.float32, seed=3)
    lora_modules.inject_lora(cpu_unet, lora_target_keys(cpu_unet), rank)
    cpu_unet.load_state_dict({k: v.detach().float().cpu() for k, v in unet.state_dict().items()})
    patch_with_oracle(cpu_unet)
    with torch.no_grad():
        want = cpu_unet(r16(x2), t, r16(ctx), cross_attention_kwargs={"scale": s2}).sample
        base = cpu_unet(r16(x2), t, r16(ctx), cross_attention_kwargs={"scale": torch.zeros_like(s2)}).sample
    rel = ((got.float().cpu() - want).norm() / want.norm()).item()
    assert rel < 3e-2, rel
    # the watermark branch is visible above that noise floor, and the two CFG halves (same latents, different context) differ
    assert ((want - base).norm() / want.norm()).item() > rel
    assert not torch.equal(got[:B], got[B:])


def test_deferred_weight_gradients_match_immediate(lm, cuda_device, monkeypatch):
    """Layers that accumulate into a flat gradient buffer queue their dUp / dDn contractions and flush them in one launch at the end
    of the backward pass (aq_lora_wgrad_batch); the result equals launching them layer by layer."""
    dev = cuda_device
    g = torch.Generator().manual_seed(3)

    def run(batch):
        monkeypatch.setattr(lm, "WGRAD_BATCH", batch)
        torch.manual_seed(0)
        layers, flats = [], []
        for (din, dout, r) in ((320, 320, 64), (320, 640, 64), (640, 320, 32), (320, 320, 128)):
            lin = lm.LoRACompatibleLinear(din, dout).to(dev, torch.bfloat16)
            lin.requires_grad_(False)
            lora = lm.LoRALinearLayer(din, dout, r)
            torch.nn.init.normal_(lora.up.weight, std=0.05)
            lora = lora.to(dev)
            for p in (lora.down.weight, lora.up.weight):
                buf = torch.zeros(p.numel(), device=dev)
                p._aq_grad = buf                      # direct accumulation target, as PPFTTrainer's flat buffer provides
                flats.append(buf)
            lin.set_lora_layer(lora)
            layers.append(lin)
        gg = torch.Generator().manual_seed(5)
        x = torch.randn(2, 512, 320, generator=gg).to(dev, torch.bfloat16).requires_grad_(True)
        sc = {64: (1 + 0.5 * torch.randn(2, 64, generator=gg)).to(dev), 32: (1 + 0.5 * torch.randn(2, 32, generator=gg)).to(dev),
              128: (1 + 0.5 * torch.randn(2, 128, generator=gg)).to(dev)}
        h = layers[0](x, sc[64])
        h = layers[1](h, sc[64])
        h = layers[2](h, sc[32])
        h = layers[3](h, sc[128])
        h.float().pow(2).mean().backward()
        torch.cuda.synchronize()
        assert not lm._WGRAD_QUEUE                      # flushed by the end-of-backward callback
        return [f.clone() for f in flats], x.grad.clone()

    imm, gx_imm = run(1)
    deferred, gx_def = run(12)
    assert torch.equal(gx_imm, gx_def)
    for a, b in zip(imm, deferred):
        assert a.abs().sum() > 0
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6 * float(a.abs().max()))       # fp32 atomics: order only
