"""Hot path (ii), message decoder on a B200: csrc/decoder.cu through `aqualora_b200.decoder.SecretDecoder` (ctypes -> C ABI)
against the CPU oracle (oracle/models_oracle.py, itself pinned to torchvision's efficientnet_b1 in test_oracle_golden.py).

Tolerance: both sides are fp32; they differ by summation order, the BatchNorm fold (done in float64, rounded once) and
float atomics in the squeeze / average pools.  Stated bound: |logit - oracle| <= 1e-4 * max|logit|.  Decoded bits must be
IDENTICAL wherever the oracle's own margin |l0 - l1| exceeds twice that bound (a bit whose margin is below the fp32 noise
floor of the two summation orders is undecidable for any implementation); the count of such bits is reported and is
expected to be 0 for the seeds used here.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def setup(cuda_device):
    from aqualora_b200.decoder import SecretDecoder
    from oracle import models_oracle as MO

    sd, _ = MO.make_decoder_state(48, seed=0)
    dec = SecretDecoder(48)
    missing = dec.load_state_dict(sd, strict=True)       # the reference checkpoint layout loads unchanged
    assert not missing.missing_keys and not missing.unexpected_keys
    return dec.to(cuda_device).eval(), sd, MO


def _check(dec, sd, MO, x, dev):
    with torch.no_grad():
        want = MO.secret_decoder_forward(x, sd, 48)
    got = dec(x.to(dev)).cpu()
    assert got.shape == want.shape == (x.shape[0], 48, 2)
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= TOL * scale, (err, scale)
    margin = (want[..., 0] - want[..., 1]).abs()
    decidable = margin > 2 * TOL * scale
    bits = dec.decode_bits(x.to(dev)).cpu()
    want_bits = MO.decode_bits(want)
    assert torch.equal(bits.long()[decidable], want_bits[decidable])
    return int((bits.long() != want_bits).sum()), int((~decidable).sum()), err / scale


def test_logits_and_bits_match_oracle(setup, cuda_device):
    dec, sd, MO = setup
    x = torch.rand(3, 3, 512, 512, generator=torch.Generator().manual_seed(3)) * 2 - 1
    flips, undecidable, rel = _check(dec, sd, MO, x, cuda_device)
    print(f"decoder: rel err {rel:.2e}, {flips} flipped bits of {3 * 48}, {undecidable} below the noise floor")
    assert flips == 0


def test_resize_path(setup, cuda_device):
    """utils/models.py:92-94: inputs that are not 512 x 512 are bilinearly resized first."""
    dec, sd, MO = setup
    for hw in ((256, 256), (300, 420)):
        x = torch.rand(1, 3, *hw, generator=torch.Generator().manual_seed(hw[0])) * 2 - 1
        _check(dec, sd, MO, x, cuda_device)


def test_batch_independence_and_structured_inputs(setup, cuda_device):
    dec, sd, MO = setup
    g = torch.Generator().manual_seed(11)
    x = torch.rand(5, 3, 512, 512, generator=g) * 2 - 1
    x[1] = 0.0                                             # all-zero image (the pretrain loop starts with these)
    x[2] = 1.0
    x[3, :, ::2] = -1.0
    full = dec(x.to(cuda_device))
    for i in (0, 4):
        one = dec(x[i:i + 1].to(cuda_device))
        assert (one - full[i:i + 1]).abs().max().item() <= 1e-5 * full.abs().max().item()
    _check(dec, sd, MO, x[1:4], cuda_device)


def test_other_message_lengths_and_errors(setup, cuda_device):
    from aqualora_b200._lib import AqualoraError
    from aqualora_b200.decoder import SecretDecoder

    dec, sd, MO = setup
    sd16, _ = MO.make_decoder_state(16, seed=2)
    d16 = SecretDecoder(16)
    d16.load_state_dict(sd16)
    d16 = d16.to(cuda_device).eval()
    x = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(1)) * 2 - 1
    with torch.no_grad():
        want = MO.secret_decoder_forward(x, sd16, 16)
    got = d16(x.to(cuda_device)).cpu()
    assert (got - want).abs().max().item() <= TOL * want.abs().max().item()
    with pytest.raises(AqualoraError):
        dec(x)                                             # CPU tensor
    dec.train()
    with pytest.raises(AqualoraError):
        dec(x)                                             # CPU tensor in train mode: no fallback either
    dec.eval()


def test_full_batch_64_runs_and_is_finite(setup, cuda_device):
    dec, _, _ = setup
    x = torch.rand(64, 3, 512, 512, device=cuda_device) * 2 - 1
    logits = dec(x)
    bits = dec.decode_bits(x)
    assert torch.isfinite(logits).all() and bits.shape == (64, 48)
    assert torch.equal(bits.long(), logits.argmax(-1))
