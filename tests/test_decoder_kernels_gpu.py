"""Kernel-level parity of the decoder's MBConv building blocks on a B200, through the C ABI (ctypes):
aq_conv1x1_tf32x3 (csrc/decoder_pw.cu: tcgen05 kind::tf32, 3-term split, fp32 accumulate in TMEM) and aq_depthwise_silu
(csrc/decoder.cu) against a float64 PyTorch evaluation of the same op (torchvision MBConv pieces used by
utils/models.py:88-96).

Tolerance: the split product drops only a_lo * w_lo (<= 2^-22 relative per term) and accumulates in fp32, so the result must
be as close to the float64 value as an fp32 accumulation chain is: |err| <= 2e-6 * max(1, sqrt(K / 256)) * (sum_k |x||w| + |b|)
element-wise (fp32 accumulation error grows like sqrt(K); measured 3e-7 ... 2.8e-6 at K = 1920).  A plain TF32 product misses
this bound by two to three orders of magnitude (checked below).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _silu(v):
    return v / (1 + torch.exp(-v))


CASES = [  # (M, K, N, hw, epi, use_se)  -- the decoder's real (K, N) pairs, incl. K < 32, N not a multiple of 16, N > 256
    (1024, 16, 96, 256, 1, False),
    (512, 32, 16, 256, 0, False),
    (512, 16, 16, 256, 2, True),
    (768, 96, 24, 256, 0, True),
    (640, 144, 40, 128, 0, True),
    (1024, 40, 240, 256, 1, False),
    (512, 240, 80, 256, 2, True),
    (256, 672, 192, 256, 0, True),
    (256, 192, 1152, 256, 1, False),
    (512, 1920, 320, 256, 2, True),
    (512, 320, 1280, 256, 3, False),
    (37 * 128, 112, 672, 128, 1, False),
]


@pytest.mark.parametrize("M,K,N,hw,epi,use_se", CASES)
def test_conv1x1_tf32x3_matches_float64(cuda_device, M, K, N, hw, epi, use_se):
    from aqualora_b200 import ops

    g = torch.Generator().manual_seed(M + 7 * K + 13 * N)
    x = torch.randn(M, K, generator=g) * 3
    w = torch.randn(N, K, generator=g) * K ** -0.5
    b = torch.randn(N, generator=g)
    se = torch.rand(M // hw, K, generator=g) if use_se else None
    res = torch.randn(M, N, generator=g) if epi == 2 else None
    xs = x if se is None else (x * se.repeat_interleave(hw, 0))          # fp32 product, rounded once (as the reference does)
    ref = xs.double() @ w.double().t() + b.double()
    mag = xs.double().abs() @ w.double().abs().t() + b.double().abs()
    if epi in (1, 3):
        ref = _silu(ref)
    if epi == 2:
        ref = ref + res.double()
        mag = mag + res.double().abs()
    if epi == 3:
        ref = ref.view(M // hw, hw, N).sum(1)
        mag = mag.view(M // hw, hw, N).sum(1)
    dev = cuda_device
    y = ops.conv1x1_tf32x3(x.to(dev), w.to(dev), b.to(dev), None if se is None else se.to(dev),
                           None if res is None else res.to(dev), hw=hw, epi=epi).cpu().double()
    assert y.shape == ref.shape
    err = ((y - ref).abs() / mag).max().item()
    assert err <= 2e-6 * max(1.0, (K / 256) ** 0.5), err
    # what a single-pass TF32 product would give (for the record: the bound above is not vacuous)
    xt = (xs.view(torch.int32) & -8192).view(torch.float32)
    wt = (w.view(torch.int32) & -8192).view(torch.float32)
    tf32 = xt.double() @ wt.double().t() + b.double()
    if epi == 0:
        assert ((tf32 - ref).abs() / mag).max().item() > 1e-4


@pytest.mark.parametrize("B,H,C,k,stride", [(2, 64, 32, 3, 1), (2, 64, 96, 3, 2), (1, 32, 144, 5, 2), (3, 16, 480, 5, 1),
                                            (2, 16, 1152, 3, 1), (1, 33, 16, 3, 1), (1, 20, 672, 5, 2),
                                            # large maps = the round-robin unit walk: 144 channels (half-empty last block, groups not a
                                            # divisor of the units), 16 channels on 64-wide tiles with a partial tile column, a single
                                            # image whose tile columns are split into chunks (fewer units than groups)
                                            (3, 128, 144, 3, 1), (2, 72, 16, 3, 1), (1, 256, 32, 3, 1), (2, 136, 96, 5, 1)])
def test_depthwise_silu_matches_float64(cuda_device, B, H, C, k, stride):
    from aqualora_b200 import ops

    g = torch.Generator().manual_seed(B + H + C + k)
    x = torch.randn(B, H, H, C, generator=g)
    w = torch.randn(C, 1, k, k, generator=g) * 0.3
    b = torch.randn(C, generator=g)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride, (k - 1) // 2, groups=C)
    ref = _silu(ref).permute(0, 2, 3, 1)
    dev = cuda_device
    y, pooled = ops.depthwise_silu(x.to(dev), w.reshape(C, k * k).t().contiguous().to(dev), b.to(dev), k, stride)
    assert y.shape == ref.shape
    assert (y.cpu().double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    want_pool = ref.sum((1, 2))
    assert (pooled.cpu().double() - want_pool).abs().max().item() <= 1e-4 * want_pool.abs().max().item() + 1e-3


@pytest.mark.parametrize("B,H,cin,cexp,k,stride", [(2, 64, 16, 96, 3, 2), (3, 32, 24, 144, 3, 1), (2, 48, 24, 144, 5, 2), (1, 256, 16, 96, 3, 2),
                                                    (5, 16, 24, 144, 3, 1)])
def test_expand_dw_fused_matches_float64_and_the_unfused_pair(cuda_device, B, H, cin, cexp, k, stride):
    """aq_expand_dw_fused (csrc/decoder_fused.cu: expand 1x1 + SiLU -> depthwise + SiLU + squeeze sums, the expanded map stays in
    shared memory) against float64 PyTorch and against the two stand-alone kernels it replaces.  Tolerances: the expand product is
    the same 3-term TF32 split (<= 2e-6 of sum |x||w|), SiLU uses the SFU ex2 / rcp (~1e-6 relative), the depthwise stage is an
    fp32 FMA chain of k*k terms: 3e-5 of the output's largest magnitude element-wise, as for aq_depthwise_silu."""
    from aqualora_b200 import ops

    g = torch.Generator().manual_seed(B + H + cin + k)
    x = torch.randn(B, H, H, cin, generator=g) * 2
    w_e = torch.randn(cexp, cin, generator=g) * cin ** -0.5
    b_e = torch.randn(cexp, generator=g)
    w_d = torch.randn(cexp, 1, k, k, generator=g) * 0.3
    b_d = torch.randn(cexp, generator=g)
    e = _silu(x.double() @ w_e.double().t() + b_e.double())                       # [B, H, H, cexp]
    ref = torch.nn.functional.conv2d(e.permute(0, 3, 1, 2), w_d.double(), b_d.double(), stride, (k - 1) // 2, groups=cexp)
    ref = _silu(ref).permute(0, 2, 3, 1)
    dev = cuda_device
    wd_flat = w_d.reshape(cexp, k * k).t().contiguous().to(dev)
    y, pooled = ops.expand_dw_fused(x.to(dev), w_e.to(dev), b_e.to(dev), wd_flat, b_d.to(dev), k, stride)
    assert y.shape == ref.shape
    assert (y.cpu().double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()
    want_pool = ref.sum((1, 2))
    assert (pooled.cpu().double() - want_pool).abs().max().item() <= 1e-4 * want_pool.abs().max().item() + 1e-3
    # the pair of kernels it replaces
    e_dev = ops.conv1x1_tf32x3(x.to(dev).reshape(-1, cin), w_e.to(dev), b_e.to(dev), hw=H * H, epi=1).view(B, H, H, cexp)
    y2, pooled2 = ops.depthwise_silu(e_dev, wd_flat, b_d.to(dev), k, stride)
    assert (y - y2).abs().max().item() <= 1e-5 * ref.abs().max().item()
    assert (pooled - pooled2).abs().max().item() <= 1e-4 * want_pool.abs().max().item() + 1e-3


def test_expand_dw_fused_rejects_other_shapes(cuda_device):
    from aqualora_b200 import _lib, ops

    dev = cuda_device
    with pytest.raises(_lib.AqualoraError):
        ops.expand_dw_fused(torch.zeros(1, 32, 32, 40, device=dev), torch.zeros(240, 40, device=dev), torch.zeros(240, device=dev),
                            torch.zeros(9, 240, device=dev), torch.zeros(240, device=dev), 3, 2)
