"""The oracle restatements (oracle/) against golden vectors produced by the REFERENCE's own Python
(tools/gen_golden.py, run in the build container where /root/reference exists).  CPU only."""
import json
import os

import pytest
import torch

from oracle import lora_oracle as O
from oracle import models_oracle as MO
from oracle import noise_oracle as NO


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_lora_linear_forward_and_grads_match_reference(golden_dir):
    cases = _load(golden_dir, "lora_linear.pt")
    assert len(cases) >= 6
    for c in cases:
        x = c["x"].clone().requires_grad_(True)
        down = c["down"].clone().requires_grad_(True)
        up = c["up"].clone().requires_grad_(True)
        scale = c["scale"]
        if isinstance(scale, torch.Tensor):
            scale = scale.clone().requires_grad_(True)
        lora = {"down": down, "up": up, "network_alpha": c["alpha"], "rank": c["r"]}
        y = O.lora_compatible_linear_forward(x, c["w"], c["b"], lora, scale)
        assert torch.equal(y, c["y"]), c["kind"]                       # same op sequence -> bit-identical in fp32
        y.backward(c["gy"])
        assert torch.equal(x.grad, c["gx"])
        assert torch.equal(down.grad, c["g_down"])
        assert torch.equal(up.grad, c["g_up"])
        if isinstance(scale, torch.Tensor):
            assert torch.equal(scale.grad, c["g_scale"])
        # lora_layer is None branch
        assert torch.equal(O.lora_compatible_linear_forward(c["x"], c["w"], c["b"], None, scale), c["y_base"])


def test_lora_closed_form_matches_reference(golden_dir):
    """SURVEY 8(a) closed form (what the CUDA kernels implement) vs the reference outputs."""
    for c in _load(golden_dir, "lora_linear.pt"):
        a = 1.0 if c["alpha"] is None else c["alpha"] / c["r"]
        if isinstance(c["scale"], torch.Tensor):
            s, f = c["scale"], 1.0
        else:
            s, f = torch.ones(c["B"], c["r"]), c["scale"]
        y = O.closed_form_linear(c["x"], c["w"], c["b"], c["down"], c["up"], s, a * f)
        torch.testing.assert_close(y, c["y"], rtol=1e-5, atol=1e-5)
        dx, dd, du, ds = O.closed_form_linear_grads(c["x"], c["w"], c["down"], c["up"], s, c["gy"], a * f)
        torch.testing.assert_close(dx, c["gx"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(dd, c["g_down"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(du, c["g_up"], rtol=1e-4, atol=1e-5)
        if c["g_scale"] is not None:
            torch.testing.assert_close(ds, c["g_scale"], rtol=1e-4, atol=1e-5)


def test_zero_scale_is_bit_identical_to_base(golden_dir):
    c = [c for c in _load(golden_dir, "lora_linear.pt") if c["kind"] == "zero"][0]
    assert torch.equal(c["y"], c["y_base"])


def test_lora_conv1x1_matches_reference(golden_dir):
    for c in _load(golden_dir, "lora_conv1x1.pt"):
        x = c["x"].clone().requires_grad_(True)
        down = c["down"].clone().requires_grad_(True)
        up = c["up"].clone().requires_grad_(True)
        scale = c["scale"].clone().requires_grad_(True) if isinstance(c["scale"], torch.Tensor) else c["scale"]
        lora = {"down": down, "up": up, "network_alpha": c["alpha"], "rank": c["r"]}
        y = O.lora_compatible_conv_forward(x, c["w"], c["b"], lora, scale)
        assert torch.equal(y, c["y"])
        y.backward(c["gy"])
        assert torch.equal(x.grad, c["gx"]) and torch.equal(down.grad, c["g_down"]) and torch.equal(up.grad, c["g_up"])
        if isinstance(scale, torch.Tensor):
            assert torch.equal(scale.grad, c["g_scale"])


def test_mapper_matches_reference(golden_dir):
    g = _load(golden_dir, "models_small.pt")["mapper"]
    assert torch.equal(O.mapper_forward(g["msg"], g["emb"]), g["scale"])
    # KATs from SURVEY 8(c): unit row std, mean ~ 1
    torch.testing.assert_close(g["emb"].std(dim=1), torch.ones(48), rtol=1e-5, atol=1e-5)
    assert abs(g["scale"].mean().item() - 1.0) < 0.2


def test_secret_encoder_matches_reference(golden_dir):
    g = _load(golden_dir, "models_small.pt")
    assert g["encoder_zero_init_is_zero"] is True
    sd = g["encoder_state"]
    for key in ("encoder_64x64", "encoder_96x96", "encoder_40x56"):
        c = g[key]
        xo, cm = MO.secret_encoder_forward(c["x"], c["msg"], sd)
        assert torch.equal(cm, c["c"]) and torch.equal(xo, c["x_out"])


def test_decoder_restatement_matches_torchvision():
    sd, module = MO.make_decoder_state(48, seed=0)
    module.eval()
    x = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(3)) * 2 - 1
    with torch.no_grad():
        want = module(x).view(-1, 48, 2)
        got = MO.secret_decoder_forward(x, sd, 48)
    assert torch.equal(got, want)
    # 256x256 input goes through the bilinear resize of utils/models.py:92-94
    x2 = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(4)) * 2 - 1
    with torch.no_grad():
        want2 = module(torch.nn.functional.interpolate(x2, size=(512, 512), mode="bilinear")).view(-1, 48, 2)
        got2 = MO.secret_decoder_forward(x2, sd, 48)
    assert torch.equal(got2, want2)


def test_threshold_kats():
    assert MO.get_threshold(48, 1e-6) == 40          # evaluation/run_eval_base.py:25 operating point
    assert MO.get_threshold(48, 1e-3) == 35


def test_jpeg_matches_reference(golden_dir):
    for c in _load(golden_dir, "jpeg_small.pt"):
        got = NO.jpeg_mask(c["x"])
        torch.testing.assert_close(got, c["y"], rtol=0, atol=2e-5)


def test_jpeg_kats():
    m = NO.zigzag_keep_mask(25)
    assert int(m.sum()) == 25 and m[0, 0] == 1 and m[7, 7] == 0
    assert int(NO.zigzag_keep_mask(9).sum()) == 9
    d, i = NO.dct_matrices(torch.float64)
    torch.testing.assert_close(i @ d, torch.eye(8, dtype=torch.float64), rtol=0, atol=1e-12)
    x = torch.rand(1, 3, 16, 24, generator=torch.Generator().manual_seed(0)) * 2 - 1
    y = NO.jpeg_mask(x, keep=(64, 64, 64))
    assert (y - x).abs().max().item() < 5e-5            # keep-all round trip (colour matrices are not exact inverses)


def test_unet_keys_golden(golden_dir):
    from aqualora_b200.unet import UNet2DConditionModel, UNetConfig, lora_target_keys

    want = json.load(open(os.path.join(golden_dir, "unet_keys.json")))
    assert len(want) == 192
    for cfg in (UNetConfig.sd15(), UNetConfig.sd21()):
        with torch.device("meta"):
            unet = UNet2DConditionModel(cfg)
        assert lora_target_keys(unet) == want


def test_create_wm_lora_oracle_matches_reference_script(golden_dir):
    """oracle/deploy_oracle.fold_message == scripts/create_wm_lora.py:create_watermark_lora (run by tools/gen_golden.py), bit for bit."""
    from oracle import deploy_oracle as DO

    g = _load(golden_dir, "create_wm_lora.pt")
    out = DO.fold_message(g["lora_sd"], g["emb"], g["hidinfo"], g["scale"])
    assert set(out) == set(g["out"]) and not any("text_encoder" in k for k in out)
    for k, v in g["out"].items():
        assert torch.equal(out[k], v), k


def test_merge_delta_oracle_known_answers():
    """scripts/merge_lora.py:98-120: rank-1 known answer, alpha scaling, 1x1-conv branch == linear branch."""
    from oracle import deploy_oracle as DO

    w = torch.zeros(3, 2)
    up = torch.tensor([[1.0], [2.0], [3.0]])
    down = torch.tensor([[10.0, 100.0]])
    assert torch.equal(DO.merge_delta(w, up, down, 0.5), torch.tensor([[5.0, 50.0], [10.0, 100.0], [15.0, 150.0]]))
    assert torch.equal(DO.merge_delta(w, up, down, 1.0, alpha=2.0), 2 * DO.merge_delta(w, up, down, 1.0))
    g = torch.Generator().manual_seed(0)
    w, up, down = torch.randn(6, 5, generator=g), torch.randn(6, 4, generator=g), torch.randn(4, 5, generator=g)
    lin = DO.merge_delta(w, up, down, 0.7)
    conv = DO.merge_delta(w[:, :, None, None], up[:, :, None, None], down[:, :, None, None], 0.7)
    assert torch.equal(conv[:, :, 0, 0], lin)
