"""SURVEY.md 8(f1) on a B200: scripts/create_wm_lora.py and scripts/merge_lora.py (this repository's drop-ins, CUDA arithmetic
through the C ABI: aq_mapper_fwd, aq_lora_fold_down, aq_lora_merge) against the golden output of the REFERENCE's own
create_watermark_lora (tests/golden/create_wm_lora.pt) and the oracle restatement of merge_lora.py:98-120.

Bars: given the same mapper output m, the folded `down` weights are BIT-EXACT (two fp32 multiplies in the reference's order).
End to end the script also evaluates m = mapper(msg) on the GPU, whose 48-term fp32 sum runs in a different order than
torch's CPU `sum(dim=1)`: |m - m_ref| <= 4 ulp, so the folded weights are within 1e-6 relative of the reference's.  The merge
is a rank-r fp32 GEMM whose summation order differs from torch's: |err| <= 1e-6 * sum|u||d|.
"""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _script(name):
    spec = importlib.util.spec_from_file_location(f"aq_script_{name}", os.path.join(ROOT, "scripts", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_create_watermark_lora_is_bit_exact_vs_reference(cuda_device, golden_dir, tmp_path):
    from safetensors.torch import load_file, save_file

    g = torch.load(os.path.join(golden_dir, "create_wm_lora.pt"), weights_only=False)
    save_file(g["lora_sd"], str(tmp_path / "pytorch_lora_weights.safetensors"))
    torch.save({"bit_embeddings.weight": g["emb"]}, tmp_path / "mapper.pt")
    cwl = _script("create_wm_lora")
    bits, out = cwl.create_watermark_lora(str(tmp_path), g["scale"], 48, g["hidinfo"], save=True)
    assert bits == g["hidinfo"]
    assert set(out) == set(g["out"])
    for k, v in g["out"].items():
        if "up.weight" in k:
            assert torch.equal(out[k], v), k
        else:
            assert (out[k] - v).abs().max().item() <= 1e-6 * v.abs().max().item(), k
    saved = load_file(str(tmp_path / bits / "pytorch_lora_weights.safetensors"))
    assert set(saved) == set(g["out"]) and all(torch.equal(saved[k], out[k]) for k in saved)


def test_fold_kernel_is_bit_exact_given_the_mapper_output(cuda_device, golden_dir):
    """aq_lora_fold_down alone, fed the reference's own m = mapper(msg): identical bits for linear and conv targets."""
    from aqualora_b200 import ops
    from oracle import deploy_oracle as DO

    g = torch.load(os.path.join(golden_dir, "create_wm_lora.pt"), weights_only=False)
    msg = torch.tensor([int(c) for c in g["hidinfo"]]).unsqueeze(0).float()
    m = DO.mapper_forward(g["emb"], msg)[0].contiguous()
    n = 0
    for k, v in g["out"].items():
        if "down.weight" in k:
            got = ops.lora_fold_down(g["lora_sd"][k].to(cuda_device), m.to(cuda_device), g["scale"]).cpu()
            assert got.shape == v.shape and torch.equal(got, v), k
            n += 1
    assert n == 3


def test_create_watermark_lora_other_ranks_and_linear_proj(cuda_device, tmp_path):
    """rank is read from mapper.pt (the reference hard-codes 320) and proj_in may be a linear layer (SD 2.x)."""
    from safetensors.torch import save_file

    from oracle import deploy_oracle as DO

    g = torch.Generator().manual_seed(3)
    r = 64
    sd = {"unet.mid_block.attentions.0.proj_in.lora.down.weight": torch.randn(r, 40, generator=g),
          "unet.mid_block.attentions.0.proj_in.lora.up.weight": torch.randn(40, r, generator=g),
          "unet.mid_block.attentions.0.transformer_blocks.0.attn2.processor.to_k_lora.down.weight": torch.randn(r, 24, generator=g),
          "unet.mid_block.attentions.0.transformer_blocks.0.attn2.processor.to_k_lora.up.weight": torch.randn(16, r, generator=g)}
    emb = torch.randn(48, r, generator=g)
    save_file(sd, str(tmp_path / "pytorch_lora_weights.safetensors"))
    torch.save({"bit_embeddings.weight": emb}, tmp_path / "mapper.pt")
    hid = "01" * 24
    _, out = _script("create_wm_lora").create_watermark_lora(str(tmp_path), 1.03, 48, hid, save=False)
    want = DO.fold_message(sd, emb, hid, 1.03)
    for k in want:
        assert (out[k] - want[k]).abs().max().item() <= 1e-6 * want[k].abs().max().item(), k


def test_merge_to_sd_model_matches_oracle(cuda_device):
    from aqualora_b200.unet import UNet2DConditionModel, UNetConfig
    from oracle import deploy_oracle as DO

    torch.manual_seed(0)
    unet = UNet2DConditionModel(UNetConfig.tiny(16)).float()
    ml = _script("merge_lora")
    g = torch.Generator().manual_seed(1)
    targets = {"down_blocks.0.attentions.0.proj_in": None, "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q": None,
               "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj": 16.0, "up_blocks.1.attentions.2.transformer_blocks.0.ff.net.2": None}
    lora_sd, want, mags = {}, {}, {}
    ranks = {"down_blocks.0.attentions.0.proj_in": 320, "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj": 100}   # the reference's
    for path, alpha in targets.items():                                       # released rank (create_wm_lora.py:19) and a ragged one
        r = ranks.get(path, 32)
        mod = unet
        for part in path.split("."):
            mod = getattr(mod, part)
        w = mod.weight.data.clone()
        dout, din = w.shape[0], w.shape[1]
        conv = w.dim() == 4
        up = torch.randn(dout, r, generator=g) * 0.1
        down = torch.randn(r, din, generator=g) * 0.1
        name = "lora_unet_" + path.replace(".", "_")
        lora_sd[name + ".lora_down.weight"] = down[:, :, None, None] if conv else down
        lora_sd[name + ".lora_up.weight"] = up[:, :, None, None] if conv else up
        if alpha is not None:
            lora_sd[name + ".alpha"] = torch.tensor(alpha)
        want[path] = DO.merge_delta(w.double(), lora_sd[name + ".lora_up.weight"].double(), lora_sd[name + ".lora_down.weight"].double(), 0.8, alpha)
        mags[path] = (up.abs().double() @ down.abs().double()).max().item() + w.abs().max().item()
    untouched = unet.conv_in.weight.data.clone()
    ml.merge_to_sd_model(None, unet, [lora_sd], [0.8], torch.float32)
    for path in targets:
        mod = unet
        for part in path.split("."):
            mod = getattr(mod, part)
        err = (mod.weight.data.double().cpu() - want[path]).abs().max().item()
        assert err <= 1e-6 * mags[path], (path, err)
    assert torch.equal(unet.conv_in.weight.data, untouched)
