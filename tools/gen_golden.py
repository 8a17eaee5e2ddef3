"""Generate tests/golden/* by executing the REFERENCE's own Python (read-only, /root/reference) on seeded inputs.

Runs only in the build container (the GPU box has no /root/reference).  The reference modules are imported
unchanged; the only shims are `sys.modules` placeholders for packages that are absent here and that the hot-path
functions never call (`diffusers`, `lpips`, `timm` -- see SURVEY.md 8(c)).

    python tools/gen_golden.py            # rewrites tests/golden/*.pt
"""
from __future__ import annotations

import importlib.util
import json
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _stub_modules():
    d = types.ModuleType("diffusers")
    dl = types.ModuleType("diffusers.loaders")
    dm = types.ModuleType("diffusers.models")
    dml = types.ModuleType("diffusers.models.lora")

    class LoraLoaderMixin:  # placeholder base class, never exercised
        pass

    class LoRACompatibleLinear(nn.Linear):
        def __init__(self, *a, lora_layer=None, **k):
            super().__init__(*a, **k)
            self.lora_layer = lora_layer

    class LoRACompatibleConv(nn.Conv2d):
        def __init__(self, *a, lora_layer=None, **k):
            super().__init__(*a, **k)
            self.lora_layer = lora_layer

    class LoRALinearLayer(nn.Module):
        def __init__(self, in_features, out_features, rank=4, network_alpha=None):
            super().__init__()
            self.down = nn.Linear(in_features, rank, bias=False)
            self.up = nn.Linear(rank, out_features, bias=False)
            self.network_alpha = network_alpha
            self.rank = rank

    class LoRAConv2dLayer(nn.Module):
        def __init__(self, in_features, out_features, rank=4, kernel_size=(1, 1), stride=(1, 1), padding=0, network_alpha=None):
            super().__init__()
            self.down = nn.Conv2d(in_features, rank, kernel_size=kernel_size, stride=stride, padding=padding, bias=False)
            self.up = nn.Conv2d(rank, out_features, kernel_size=(1, 1), stride=(1, 1), bias=False)
            self.network_alpha = network_alpha
            self.rank = rank

    dl.LoraLoaderMixin = LoraLoaderMixin
    dml.text_encoder_attn_modules = lambda *a, **k: []
    dml.text_encoder_mlp_modules = lambda *a, **k: []
    dml.PatchedLoraProjection = type("PatchedLoraProjection", (nn.Module,), {})
    dml.LoRALinearLayer = LoRALinearLayer
    dml.LoRAConv2dLayer = LoRAConv2dLayer
    dml.LoRACompatibleConv = LoRACompatibleConv
    dml.LoRACompatibleLinear = LoRACompatibleLinear
    d.loaders = dl
    d.models = dm
    dm.lora = dml
    sys.modules.update({"diffusers": d, "diffusers.loaders": dl, "diffusers.models": dm, "diffusers.models.lora": dml})
    for name in ("lpips", "timm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    return dml


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_lora(dml, ref_lora):
    import types as _t

    g = torch.Generator().manual_seed(20240517)
    cases = []
    # (B, N, din, dout, r, bias, alpha, scale_kind)
    specs = [
        (2, 24, 32, 48, 8, True, None, "tensor"),
        (3, 16, 64, 32, 16, False, None, "tensor"),
        (2, 10, 40, 40, 8, True, 4.0, "tensor"),
        (2, 12, 32, 64, 8, True, None, "float"),
        (2, 12, 32, 64, 8, True, 2.0, "float"),
        (2, 77, 96, 40, 64, False, None, "tensor"),
        (2, 8, 32, 32, 8, True, None, "zero"),
    ]
    for (B, N, din, dout, r, bias, alpha, kind) in specs:
        lin = dml.LoRACompatibleLinear(din, dout, bias=bias)
        lora = dml.LoRALinearLayer(din, dout, rank=r, network_alpha=alpha)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(dout, din, generator=g) * din ** -0.5)
            if bias:
                lin.bias.copy_(torch.randn(dout, generator=g) * 0.1)
            lora.down.weight.copy_(torch.randn(r, din, generator=g) / r)
            lora.up.weight.copy_(torch.randn(dout, r, generator=g) * 0.05)
        lin.lora_layer = lora
        lin.forward = _t.MethodType(ref_lora.CustomLoRACompatibleLinearforward, lin)
        lora.forward = _t.MethodType(ref_lora.CustomLoRALinearLayerforward, lora)
        x = torch.randn(B, N, din, generator=g, requires_grad=True)
        if kind == "tensor":
            scale = (1 + 0.7 * torch.randn(B, r, generator=g)).requires_grad_(True)
        elif kind == "zero":
            scale = torch.zeros(B, r, requires_grad=True)
        else:
            scale = 0.75
        y = lin(x, scale)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        rec = {
            "B": B, "N": N, "din": din, "dout": dout, "r": r, "alpha": alpha, "kind": kind,
            "x": x.detach().clone(), "w": lin.weight.detach().clone(), "b": None if not bias else lin.bias.detach().clone(),
            "down": lora.down.weight.detach().clone(), "up": lora.up.weight.detach().clone(),
            "scale": scale.detach().clone() if isinstance(scale, torch.Tensor) else scale,
            "gy": gy, "y": y.detach().clone(), "gx": x.grad.clone(), "g_down": lora.down.weight.grad.clone(),
            "g_up": lora.up.weight.grad.clone(),
            "g_scale": scale.grad.clone() if isinstance(scale, torch.Tensor) else None,
        }
        # the lora_layer is None branch (lora_modules.py:57-59)
        lin.lora_layer = None
        rec["y_base"] = lin(x.detach(), scale).detach().clone()
        cases.append(rec)
    torch.save(cases, OUT / "lora_linear.pt")

    # conv 1x1 (proj_in / proj_out of SD1.5: lora_modules.py:28-54)
    conv_cases = []
    for (B, C, Co, H, W, r, alpha, kind) in [(2, 32, 32, 8, 8, 8, None, "tensor"), (2, 48, 24, 6, 10, 16, 8.0, "tensor"),
                                             (1, 32, 32, 4, 4, 8, None, "float")]:
        conv = dml.LoRACompatibleConv(C, Co, kernel_size=1)
        lora = dml.LoRAConv2dLayer(C, Co, rank=r, kernel_size=(1, 1), network_alpha=alpha)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(Co, C, 1, 1, generator=g) * C ** -0.5)
            conv.bias.copy_(torch.randn(Co, generator=g) * 0.1)
            lora.down.weight.copy_(torch.randn(r, C, 1, 1, generator=g) / r)
            lora.up.weight.copy_(torch.randn(Co, r, 1, 1, generator=g) * 0.05)
        conv.lora_layer = lora
        conv.forward = _t.MethodType(ref_lora.CustomLoRACompatibleConvforward, conv)
        lora.forward = _t.MethodType(ref_lora.CustomLoRAConv2dLayerforward, lora)
        x = torch.randn(B, C, H, W, generator=g, requires_grad=True)
        scale = (1 + 0.7 * torch.randn(B, r, generator=g)).requires_grad_(True) if kind == "tensor" else 1.25
        y = conv(x, scale)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        conv_cases.append({
            "x": x.detach().clone(), "w": conv.weight.detach().clone(), "b": conv.bias.detach().clone(),
            "down": lora.down.weight.detach().clone(), "up": lora.up.weight.detach().clone(), "alpha": alpha, "r": r,
            "scale": scale.detach().clone() if isinstance(scale, torch.Tensor) else scale, "gy": gy, "y": y.detach().clone(),
            "gx": x.grad.clone(), "g_down": lora.down.weight.grad.clone(), "g_up": lora.up.weight.grad.clone(),
            "g_scale": scale.grad.clone() if isinstance(scale, torch.Tensor) else None,
        })
    torch.save(conv_cases, OUT / "lora_conv1x1.pt")


def gen_models(ref_models):
    g = torch.Generator().manual_seed(7)
    torch.manual_seed(11)
    out = {}
    # MapperNet (utils/models.py:98-115)
    mapper = ref_models.MapperNet(input_size=48, output_size=64)
    msg = torch.randint(0, 2, (5, 48), generator=g).float()
    out["mapper"] = {"emb": mapper.bit_embeddings.weight.detach().clone(), "msg": msg, "scale": mapper(msg).detach().clone()}
    # SecretEncoder (utils/models.py:51-81); the conv is zero-initialised, re-initialise it so the output is informative
    enc = ref_models.SecretEncoder(48)
    with torch.no_grad():
        enc.secret_scaler[-1].weight.normal_(0, 0.02, generator=g)
        enc.secret_scaler[-1].bias.normal_(0, 0.02, generator=g)
    for hw in ((64, 64), (96, 96), (40, 56)):
        x = torch.randn(2, 4, *hw, generator=g)
        m = torch.randint(0, 2, (2, 48), generator=g).float()
        xo, c = enc(x, m)
        out[f"encoder_{hw[0]}x{hw[1]}"] = {"x": x, "msg": m, "x_out": xo.detach().clone(), "c": c.detach().clone()}
    out["encoder_state"] = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    enc0 = ref_models.SecretEncoder(48)
    out["encoder_zero_init_is_zero"] = bool((enc0(torch.zeros(1, 4, 64, 64), torch.ones(1, 48))[1] == 0).all())
    torch.save(out, OUT / "models_small.pt")


def gen_jpeg():
    sys.path.insert(0, str(REF))
    from utils.noise_layers.jpeg_compression import JpegCompression  # imports with no stubs

    g = torch.Generator().manual_seed(99)
    jp = JpegCompression("cpu")
    out = []
    for shape in ((2, 3, 64, 64), (1, 3, 40, 72), (1, 3, 37, 50)):
        x = torch.rand(shape, generator=g) * 2 - 1
        y = jp([x.clone(), None])[0]
        out.append({"x": x, "y": y.clone()})
    torch.save(out, OUT / "jpeg_small.pt")


def gen_keys():
    keys = json.load(open(REF / "utils" / "unet_keys.json"))
    (OUT / "unet_keys.json").write_text(json.dumps(keys, indent=0))


def gen_create_wm_lora():
    """scripts/create_wm_lora.py run unchanged on a synthetic train folder (rank 320 as the reference hard-codes, tiny widths)."""
    import tempfile

    from safetensors.torch import save_file

    g = torch.Generator().manual_seed(21)
    r = 320
    sd = {
        "unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.to_q_lora.down.weight": torch.randn(r, 24, generator=g),
        "unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.to_q_lora.up.weight": torch.randn(24, r, generator=g),
        "unet.mid_block.attentions.0.transformer_blocks.0.ff.net.2.lora.down.weight": torch.randn(r, 40, generator=g),
        "unet.mid_block.attentions.0.transformer_blocks.0.ff.net.2.lora.up.weight": torch.randn(16, r, generator=g),
        "unet.up_blocks.1.attentions.2.proj_in.lora.down.weight": torch.randn(r, 8, 1, 1, generator=g),
        "unet.up_blocks.1.attentions.2.proj_in.lora.up.weight": torch.randn(8, r, 1, 1, generator=g),
        "text_encoder.foo.lora.down.weight": torch.randn(4, 4, generator=g),
    }
    emb = torch.randn(48, r, generator=g)
    hid = "".join(str(int(b)) for b in torch.randint(0, 2, (48,), generator=g))
    with tempfile.TemporaryDirectory() as td:
        save_file(sd, f"{td}/pytorch_lora_weights.safetensors")
        torch.save({"bit_embeddings.weight": emb}, f"{td}/mapper.pt")
        sys.path.insert(0, str(REF / "scripts"))
        sys.path.insert(0, str(REF))
        for m in ("lpips", "timm"):
            sys.modules.setdefault(m, types.ModuleType(m))
        cwl = _load("ref_create_wm_lora", REF / "scripts" / "create_wm_lora.py")
        bits, out = cwl.create_watermark_lora(td, 1.03, 48, hid, save=False)
    assert bits == hid
    torch.save({"lora_sd": sd, "emb": emb, "hidinfo": hid, "scale": 1.03, "out": {k: v.detach().clone() for k, v in out.items()}},
               OUT / "create_wm_lora.pt")


def gen_pretrain():
    """PRVL_loss and gen_combined_latents of train/latent_wm_pretrain.py, executed from the reference's own source text: the
    script cannot be imported (accelerate, diffusers, lpips, torchsummary are absent), so the two function bodies are cut out by
    line range, checked by their first lines, and exec'd."""
    import random
    import textwrap

    import torch.nn.functional as F

    lines = (REF / "train" / "latent_wm_pretrain.py").read_text().splitlines()
    prvl_src = "\n".join(lines[38:50])                       # WINDOW_SIZE, KERNEL, def PRVL_loss
    assert prvl_src.startswith("WINDOW_SIZE = 32") and "def PRVL_loss(img1, img2):" in prvl_src
    ns = {"torch": torch, "F": F}
    exec(prvl_src, ns)
    comb_src = textwrap.dedent("\n".join(lines[132:150]))    # def gen_combined_latents (a closure of main(): uses random, F, torch)
    assert comb_src.startswith("def gen_combined_latents(latents, wm_latent, scale=1.0):")
    ns2 = {"torch": torch, "F": F, "random": random}
    exec(comb_src, ns2)
    g = torch.Generator().manual_seed(0)
    out = {"prvl": [], "combined": []}
    for shape in [(2, 3, 64, 64), (1, 3, 96, 80), (2, 3, 512, 512)]:
        a = torch.rand(shape, generator=g) * 2 - 1
        b = a + 0.1 * torch.randn(shape, generator=g)
        keep = shape[-1] <= 96
        out["prvl"].append({"shape": shape, "a": a if keep else None, "b": b if keep else None, "seed_note": "512: regenerate from seed",
                            "value": ns["PRVL_loss"](a, b).clone()})
    out["prvl_512_inputs_seed"] = 0
    for seed in range(8):
        lat = torch.randn(2, 4, 16, 16, generator=g)
        wm = torch.randn(2, 4, 16, 16, generator=g) * 0.1
        random.seed(seed)
        y = ns2["gen_combined_latents"](lat.clone(), wm.clone(), scale=0.03 if seed % 2 else 1.0)
        out["combined"].append({"seed": seed, "scale": 0.03 if seed % 2 else 1.0, "latents": lat, "wm": wm, "out": y.clone()})
    torch.save(out, OUT / "pretrain_small.pt")


def _ref_unet_module():
    return _load("ref_original_unet", REF / "scripts" / "lib" / "original_unet.py")


def _unet_cfgs():
    return {
        "sd15": dict(sample_size=64, attention_head_dim=8, cross_attention_dim=768),
        "sd21": dict(sample_size=96, attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024, use_linear_projection=True,
                     upcast_attention=True),
    }


def gen_unet():
    """The reference's vendored U-Net (scripts/lib/original_unet.py:1311-1585, full SD1.5 / SD2.1 widths) on procedurally
    generated weights (tests/procedural.py: seeded per state-dict key, so the 3.4 GB state dict is never stored) at small latent
    sizes.  tests/test_unet_golden.py loads the SAME tensors into aqualora_b200.unet and compares outputs."""
    sys.path.insert(0, str(OUT.parent))
    from procedural import load_procedural

    m = _ref_unet_module()
    g = torch.Generator().manual_seed(4242)
    out = {}
    for name, kw in _unet_cfgs().items():
        with torch.device("meta"):
            unet = m.UNet2DConditionModel(**kw)
        load_procedural(unet, seed=0)
        unet.eval()
        cases = []
        specs = [(2, 16, 16, [17, 803]), (1, 20, 12, [500])] if name == "sd15" else [(1, 24, 24, [999])]
        for (B, H, W, ts) in specs:
            x = torch.randn(B, 4, H, W, generator=g)
            ctx = torch.randn(B, 77, kw["cross_attention_dim"], generator=g)
            t = torch.tensor(ts)
            with torch.no_grad():
                y = unet(x, t, ctx).sample
            cases.append({"x": x, "t": t, "ctx": ctx, "y": y.clone()})
        out[name] = {"kwargs": kw, "seed": 0, "n_params": sum(p.numel() for p in unet.parameters()), "cases": cases}
        del unet
    torch.save(out, OUT / "unet_reference.pt")


def gen_unet_lora_step(dml, ref_lora):
    """One PPFT step body (train/ppft_train.py:1026-1058: clean forward with an all-zero scale, watermarked forward with the
    mapped scale, MSE, backward) on the reference's vendored SD1.5 U-Net whose 192 unet_keys.json targets were swapped for
    LoRACompatible{Linear,Conv} containers and monkey-patched with the reference's own utils/lora_modules.py forwards
    (train/ppft_train.py:681-689).  The vendored U-Net calls its projections without `scale`, so the per-call tensor scale that
    diffusers threads through cross_attention_kwargs is bound into each patched forward instead.  fp32, CPU, rank 64."""
    import types as _t

    sys.path.insert(0, str(OUT.parent))
    from procedural import load_procedural, procedural_tensor

    m = _ref_unet_module()
    kw = _unet_cfgs()["sd15"]
    with torch.device("meta"):
        unet = m.UNet2DConditionModel(**kw)
    load_procedural(unet, seed=0)
    unet.requires_grad_(False)
    keys = json.load(open(REF / "utils" / "unet_keys.json"))
    r, up_gain = 64, 0.2
    current = {"scale": 1.0}
    loras = []
    for key in keys:
        parts = key.split(".")
        parent = unet
        for sub in parts[:-1]:
            parent = getattr(parent, sub)
        old = getattr(parent, parts[-1]) if not parts[-1].isdigit() else parent[int(parts[-1])]
        if isinstance(old, nn.Conv2d):
            new = dml.LoRACompatibleConv(old.in_channels, old.out_channels, kernel_size=1)
            lora = dml.LoRAConv2dLayer(old.in_channels, old.out_channels, rank=r)
            lin_fwd, lora_fwd = ref_lora.CustomLoRACompatibleConvforward, ref_lora.CustomLoRAConv2dLayerforward
            shp_d, shp_u = (r, old.in_channels, 1, 1), (old.out_channels, r, 1, 1)
        else:
            new = dml.LoRACompatibleLinear(old.in_features, old.out_features, bias=old.bias is not None)
            lora = dml.LoRALinearLayer(old.in_features, old.out_features, rank=r)
            lin_fwd, lora_fwd = ref_lora.CustomLoRACompatibleLinearforward, ref_lora.CustomLoRALinearLayerforward
            shp_d, shp_u = (r, old.in_features), (old.out_features, r)
        new.weight, new.bias = old.weight, old.bias
        with torch.no_grad():
            lora.down.weight.copy_(procedural_tensor(key + ".lora_layer.down.weight", shp_d, 1))
            lora.up.weight.copy_(procedural_tensor(key + ".lora_layer.up.weight", shp_u, 1) * up_gain)
        new.lora_layer = lora
        lora.forward = _t.MethodType(lora_fwd, lora)
        new.forward = (lambda f, mod: (lambda hs: f(mod, hs, current["scale"])))(lin_fwd, new)
        if parts[-1].isdigit():
            parent[int(parts[-1])] = new
        else:
            setattr(parent, parts[-1], new)
        loras.append((key, lora))
    g = torch.Generator().manual_seed(777)
    B, H, W = 2, 16, 16
    x_clean = torch.randn(B, 4, H, W, generator=g)
    x_wm = x_clean + 0.05 * torch.randn(B, 4, H, W, generator=g)
    ctx = torch.randn(B, 77, 768, generator=g)
    t = torch.tensor([123, 871])
    scale = (1 + 0.5 * torch.randn(B, r, generator=g)).requires_grad_(True)
    current["scale"] = torch.zeros_like(scale)
    clean = unet(x_clean, t, ctx).sample.detach()
    current["scale"] = scale
    pred = unet(x_wm, t, ctx).sample
    loss = torch.nn.functional.mse_loss(pred.float(), clean.float(), reduction="mean")
    loss.backward()
    keep = {keys[0], keys[5], keys[60], keys[96], keys[107], keys[191]}
    grads, norms = {}, {}
    for key, lora in loras:
        for which, p in (("down", lora.down.weight), ("up", lora.up.weight)):
            norms[f"{key}.{which}"] = float(p.grad.norm())
            if key in keep:
                grads[f"{key}.{which}"] = p.grad.clone()
    torch.save({"rank": r, "up_gain": up_gain, "lora_seed": 1, "unet_seed": 0, "x_clean": x_clean, "x_wm": x_wm, "ctx": ctx, "t": t,
                "scale": scale.detach().clone(), "clean_pred": clean.clone(), "model_pred": pred.detach().clone(),
                "loss": float(loss), "g_scale": scale.grad.clone(), "grad_norms": norms, "grads": grads}, OUT / "unet_lora_step.pt")


def main():
    if '--pretrain-only' in sys.argv:
        gen_pretrain()
        return
    if '--create-wm-lora-only' in sys.argv:
        gen_create_wm_lora()
        return
    if '--unet-only' in sys.argv:
        dml = _stub_modules()
        gen_unet()
        gen_unet_lora_step(dml, _load("ref_lora_modules", REF / "utils" / "lora_modules.py"))
        return
    OUT.mkdir(parents=True, exist_ok=True)
    dml = _stub_modules()
    ref_lora = _load("ref_lora_modules", REF / "utils" / "lora_modules.py")
    gen_lora(dml, ref_lora)
    ref_models = _load("ref_models", REF / "utils" / "models.py")
    gen_models(ref_models)
    gen_jpeg()
    gen_keys()
    gen_create_wm_lora()
    gen_pretrain()
    gen_unet()
    gen_unet_lora_step(dml, ref_lora)
    for f in sorted(OUT.glob("*")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
