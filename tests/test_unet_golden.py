"""aqualora_b200.unet (the U-Net harness both bench arms run) against the REFERENCE's vendored U-Net
(scripts/lib/original_unet.py:1311-1585), and the oracle PPFT step composition against the reference step body
(train/ppft_train.py:1026-1058) run with the reference's own utils/lora_modules.py forwards.

The goldens (tests/golden/unet_reference.pt, unet_lora_step.pt) were produced by tools/gen_golden.py in the build
container by importing the reference unchanged; weights are procedural (tests/procedural.py), so the same tensors are
regenerated here and loaded into this repository's modules.  CPU, fp32, full SD1.5 / SD2.1 widths at 16x16 .. 24x24 latents."""
import os

import pytest
import torch

from aqualora_b200 import lora_modules
from aqualora_b200.unet import UNet2DConditionModel, UNetConfig, lora_target_keys
from oracle.patch import patch_with_oracle
from procedural import load_procedural, procedural_tensor


def _build(name):
    cfg = UNetConfig.sd15(64) if name == "sd15" else UNetConfig.sd21(96)
    with torch.device("meta"):
        unet = UNet2DConditionModel(cfg)
    load_procedural(unet, seed=0)
    unet.requires_grad_(False)
    return unet.eval()


@pytest.mark.parametrize("name", ["sd15", "sd21"])
def test_unet_matches_reference_unet(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, "unet_reference.pt"), weights_only=False)[name]
    unet = _build(name)
    patch_with_oracle(unet)          # the product's projections have no CPU path; lora_layer is None -> plain F.linear / conv2d
    assert sum(p.numel() for p in unet.parameters()) == gold["n_params"]
    for c in gold["cases"]:
        with torch.no_grad():
            y = unet(c["x"], c["t"], c["ctx"]).sample
        assert y.shape == c["y"].shape
        # same fp32 arithmetic in a different op order (fused SDPA vs baddbmm + softmax, channel-last glue): elementwise
        torch.testing.assert_close(y, c["y"], rtol=2e-4, atol=2e-4 * float(c["y"].abs().max()))


def test_oracle_ppft_step_matches_reference_step(golden_dir):
    """U-Net harness + oracle LoRA forwards == reference U-Net + reference LoRA forwards: model_pred, loss, d(scale) and the
    LoRA weight gradients of one SD1.5-width PPFT step body."""
    gold = torch.load(os.path.join(golden_dir, "unet_lora_step.pt"), weights_only=False)
    unet = _build("sd15")
    keys = lora_target_keys(unet)
    layers = lora_modules.inject_lora(unet, keys, gold["rank"])
    for key, _, lora in layers:
        with torch.no_grad():
            lora.down.weight.copy_(procedural_tensor(key + ".lora_layer.down.weight", tuple(lora.down.weight.shape), gold["lora_seed"]))
            lora.up.weight.copy_(procedural_tensor(key + ".lora_layer.up.weight", tuple(lora.up.weight.shape), gold["lora_seed"]) * gold["up_gain"])
        lora.down.weight.requires_grad_(True)
        lora.up.weight.requires_grad_(True)
    patch_with_oracle(unet)
    scale = gold["scale"].clone().requires_grad_(True)
    clean = unet(gold["x_clean"], gold["t"], gold["ctx"], cross_attention_kwargs={"scale": torch.zeros_like(scale)}).sample.detach()
    pred = unet(gold["x_wm"], gold["t"], gold["ctx"], cross_attention_kwargs={"scale": scale}).sample
    loss = torch.nn.functional.mse_loss(pred.float(), clean.float(), reduction="mean")
    loss.backward()
    amax = float(gold["model_pred"].abs().max())
    torch.testing.assert_close(clean, gold["clean_pred"], rtol=2e-4, atol=2e-4 * amax)
    torch.testing.assert_close(pred.detach(), gold["model_pred"], rtol=2e-4, atol=2e-4 * amax)
    assert abs(float(loss) - gold["loss"]) <= 1e-4 * gold["loss"]
    torch.testing.assert_close(scale.grad, gold["g_scale"], rtol=2e-3, atol=2e-3 * float(gold["g_scale"].abs().max()))
    by_key = {k: l for k, _, l in layers}
    for name, want in gold["grad_norms"].items():
        key, which = name.rsplit(".", 1)
        got = float(getattr(by_key[key], which).weight.grad.norm())
        assert abs(got - want) <= 2e-3 * want + 1e-12, (name, got, want)
    for name, want in gold["grads"].items():
        key, which = name.rsplit(".", 1)
        got = getattr(by_key[key], which).weight.grad
        torch.testing.assert_close(got, want, rtol=5e-3, atol=2e-3 * float(want.abs().max()))
