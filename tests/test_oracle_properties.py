"""Property checks of the LoRA oracle over random small shapes (hypothesis): the op-sequence restatement of
utils/lora_modules.py:9-26,56-62 (pinned bit-for-bit to the reference by tests/golden) against the closed form that the CUDA
kernels implement, its hand-derived gradients against autograd, and the size-independent identities the GPU
tests rely on at full size (zero scale = base op, linearity in up, a float scale equals a constant diagonal)."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import lora_oracle as O

dims = st.tuples(st.integers(1, 3), st.integers(1, 9), st.sampled_from([8, 24, 40]), st.sampled_from([8, 16, 56]), st.sampled_from([1, 4, 8]),
                 st.booleans(), st.sampled_from([None, 4.0]), st.integers(0, 2 ** 16))


def _case(B, N, din, dout, r, with_bias, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, din, generator=g, dtype=torch.float64)
    w = torch.randn(dout, din, generator=g, dtype=torch.float64) / din ** 0.5
    b = torch.randn(dout, generator=g, dtype=torch.float64) if with_bias else None
    dn = torch.randn(r, din, generator=g, dtype=torch.float64) / din ** 0.5
    up = torch.randn(dout, r, generator=g, dtype=torch.float64) * 0.3
    s = 1 + 0.7 * torch.randn(B, r, generator=g, dtype=torch.float64)
    gy = torch.randn(B, N, dout, generator=g, dtype=torch.float64)
    return x, w, b, dn, up, s, gy


@settings(max_examples=40, deadline=None)
@given(dims)
def test_op_sequence_equals_closed_form_and_gradients(d):
    B, N, din, dout, r, with_bias, alpha, seed = d
    x, w, b, dn, up, s, gy = _case(B, N, din, dout, r, with_bias, seed)
    a = 1.0 if alpha is None else alpha / r
    leaves = [t.clone().requires_grad_(True) for t in (x, dn, up, s)]
    xr, dnr, upr, sr = leaves
    y = O.lora_compatible_linear_forward(xr, w, b, {"down": dnr, "up": upr, "network_alpha": alpha, "rank": r}, sr)
    want = O.closed_form_linear(x, w, b, dn, up, s, a)
    torch.testing.assert_close(y, want, rtol=1e-10, atol=1e-10)
    y.backward(gy)
    dx, d_dn, d_up, d_s = O.closed_form_linear_grads(x, w, dn, up, s, gy, a)
    for got, ref in ((xr.grad, dx), (dnr.grad, d_dn), (upr.grad, d_up), (sr.grad, d_s)):
        torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-9)


@settings(max_examples=25, deadline=None)
@given(dims)
def test_identities(d):
    B, N, din, dout, r, with_bias, alpha, seed = d
    x, w, b, dn, up, s, _ = _case(B, N, din, dout, r, with_bias, seed)
    lora = {"down": dn, "up": up, "network_alpha": alpha, "rank": r}
    base = O.lora_compatible_linear_forward(x, w, b, None, 1.0)
    # zero diagonal: exactly the base op (the PPFT clean pass, train/ppft_train.py:1026-1029)
    assert torch.equal(O.lora_compatible_linear_forward(x, w, b, lora, torch.zeros_like(s)), base)
    # linear in `up`: doubling it doubles the LoRA branch
    y1 = O.lora_compatible_linear_forward(x, w, b, lora, s) - base
    y2 = O.lora_compatible_linear_forward(x, w, b, {**lora, "up": 2 * up}, s) - base
    torch.testing.assert_close(y2, 2 * y1, rtol=1e-10, atol=1e-10)
    # a float scale f (utils/lora_modules.py:24-25) equals the constant diagonal f on every sample
    f = 0.37
    yf = O.lora_compatible_linear_forward(x, w, b, lora, f)
    yd = O.lora_compatible_linear_forward(x, w, b, lora, torch.full_like(s, f))
    torch.testing.assert_close(yf, yd, rtol=1e-10, atol=1e-10)


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 4), st.sampled_from([8, 48]), st.sampled_from([4, 8, 64]), st.integers(0, 2 ** 16))
def test_mapper_is_affine_in_the_message(B, bits, r, seed):
    g = torch.Generator().manual_seed(seed)
    E = O.mapper_init(bits, r, generator=g)
    m1 = torch.randint(0, 2, (B, bits), generator=g).float()
    m2 = torch.randint(0, 2, (B, bits), generator=g).float()
    s1, s2, s12 = O.mapper_forward(m1, E), O.mapper_forward(m2, E), O.mapper_forward(m1 + m2, E)
    torch.testing.assert_close(s12 - 1, (s1 - 1) + (s2 - 1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(O.mapper_forward(torch.zeros(B, bits), E), torch.ones(B, r))
