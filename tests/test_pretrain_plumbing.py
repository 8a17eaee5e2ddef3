"""BASELINE.json configs[0]: the pretrain step of train/latent_wm_pretrain.py:164-217 with a 48-bit message, 64x64 latents and
batch 2 on the CPU (plumbing, no GPU) through the oracle restatement, plus the golden vectors that pin the two functions of that
script which are restated (PRVL_loss :42-50, gen_combined_latents :133-149; outputs of the reference's own source,
tools/gen_golden.py)."""
import os
import random

import pytest
import torch

from oracle import noise_oracle as NO
from oracle import pretrain_oracle as PO


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "pretrain_small.pt"), weights_only=False)


def test_prvl_loss_matches_reference(golden):
    g = torch.Generator().manual_seed(golden["prvl_512_inputs_seed"])
    for c in golden["prvl"]:
        a = torch.rand(c["shape"], generator=g) * 2 - 1          # same draw order as the generator script
        b = a + 0.1 * torch.randn(c["shape"], generator=g)
        if c["a"] is not None:
            assert torch.equal(a, c["a"]) and torch.equal(b, c["b"])
        assert torch.equal(PO.prvl_loss(a, b), c["value"])


def test_gen_combined_latents_matches_reference(golden):
    seen = set()
    for c in golden["combined"]:
        cornerfy, hs, ws = PO.draw_cornerfy(random.Random(c["seed"]))
        seen.add(cornerfy)
        got = PO.gen_combined_latents(c["latents"].clone(), c["wm"].clone(), c["scale"], cornerfy, hs, ws)
        assert torch.equal(got, c["out"]), c["seed"]
    assert seen == {True, False}                                  # both branches of :134-146 are covered


def test_prvl_known_answers():
    x = torch.zeros(1, 3, 64, 64)
    y = x.clone()
    y[:, :, 10:42, 10:42] = 0.5                                   # one full 32 x 32 window of |diff| = 0.5 in every channel
    assert PO.prvl_loss(x, y).item() == pytest.approx(0.5, rel=1e-6)
    assert PO.prvl_loss(x, x).item() == 0.0


@pytest.mark.parametrize("warmup,stage,layer", [(True, 0, 0), (False, 2, 1), (False, 1, 4)])
def test_config0_pretrain_step_plumbing(warmup, stage, layer):
    torch.manual_seed(0)
    B, bits = 2, 48
    enc = PO.SecretEncoderRef(bits)
    torch.nn.init.normal_(enc.secret_scaler[5].weight, std=0.02)    # past the zero init, so that every parameter sees a gradient
    dec = PO.SecretDecoderRef(bits)
    vae = PO.StubVAE(seed=0)
    enc.train(); dec.train()
    image = torch.rand(B, 3, 512, 512) * 2 - 1
    msg = torch.randint(0, 2, (B, bits))
    assert vae.encode(image).shape == (B, 4, 64, 64)
    rng_np = __import__("numpy").random.default_rng(7)
    params = NO.draw_params(rng_np, layer, B)
    noise = torch.randn(B, 3, 512, 512) if layer == 4 else None
    out = PO.pretrain_step(enc, dec, vae, image, msg, layer, params, random.Random(3), warmup, stage, noise)
    assert out["reveal"].shape == (B, bits, 2) and out["wm_image"].shape == (B, 3, 512, 512)
    for k in ("loss", "msgloss", "lpips", "prvl"):
        assert torch.isfinite(out[k]).item(), k
    assert out["msgloss"].item() == pytest.approx(0.693, abs=0.15)   # random-init logits: BCE near ln 2
    for name, p in list(enc.named_parameters()) + list(dec.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert enc.secret_scaler[0].weight.grad.abs().max() > 0 and enc.secret_scaler[5].weight.grad.abs().max() > 0
    assert dec.model.classifier[1].weight.grad.abs().max() > 0 and dec.model.features[0][0].weight.grad.abs().max() > 0
    ckpt = PO.checkpoint_dict(enc, dec)
    assert set(ckpt) == {"sec_decoder", "sec_encoder"}                # train/latent_wm_pretrain.py:246-249
    assert set(ckpt["sec_encoder"]) == {"secret_scaler.0.weight", "secret_scaler.0.bias", "secret_scaler.5.weight", "secret_scaler.5.bias"}
    assert "model.classifier.1.weight" in ckpt["sec_decoder"] and ckpt["sec_decoder"]["model.classifier.1.weight"].shape == (2 * bits, 1280)


def test_pretrained_checkpoint_loads_into_the_cuda_modules_state_dict_layout():
    """The checkpoint written by the pretrain step is what train/ppft_train.py:550-554 loads: the drop-in SecretEncoder /
    SecretDecoder of this repository must accept it key for key (CPU: state-dict plumbing only, no kernels run)."""
    from aqualora_b200.decoder import SecretDecoder
    from aqualora_b200.models import SecretEncoder

    ckpt = PO.checkpoint_dict(PO.SecretEncoderRef(48), PO.SecretDecoderRef(48))
    enc, dec = SecretEncoder(48), SecretDecoder(48)
    assert enc.load_state_dict(ckpt["sec_encoder"], strict=True) is not None
    assert dec.load_state_dict(ckpt["sec_decoder"], strict=True) is not None
