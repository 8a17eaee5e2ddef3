"""Drop-in for the merge step of the reference's `scripts/merge_lora.py` (kohya-style): W <- W + ratio * (up @ down) * alpha / dim.

`merge_to_sd_model(text_encoder, unet, models, ratios, merge_dtype)` keeps the reference's signature and its LoRA naming
(merge_lora.py:56-127): module `down_blocks.0.attentions.0.proj_in` <-> `lora_unet_down_blocks_0_attentions_0_proj_in.lora_down.weight`,
every Linear / Conv2d under a `Transformer2DModel`.  The rank-r update of every targeted weight runs on the B200
(`aq_lora_merge`, csrc/lora_deploy.cu); linear and 1x1-conv targets are implemented (all 192 `utils/unet_keys.json` targets), 3x3 LoRA
convolutions and text-encoder LoRA are not part of AquaLoRA's training (`--train_text_encoder` off) and raise.
Checkpoint I/O (`--sd_model`, `--save_to`, LDM <-> diffusers key conversion) is the reference's `scripts/lib/model_util.py`
and stays there; `merge(args)` below operates on DIFFUSERS-KEYED U-NET STATE DICTS (`.safetensors` / `.pt`), not on full SD
`.ckpt` files: the output holds the U-Net tensors only (no text encoder / VAE / sai metadata).  To merge into a full checkpoint
keep the reference's `merge(args)` and swap only its `merge_to_sd_model` for the one below.  Any rank is
supported (the released LoRAs are rank 320).
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from safetensors.torch import load_file, save_file  # noqa: E402

from aqualora_b200 import ops  # noqa: E402

LORA_PREFIX_UNET = "lora_unet"
UNET_TARGET_REPLACE_MODULE = ["Transformer2DModel"]


def load_state_dict(file_name, dtype):
    sd = load_file(file_name) if os.path.splitext(file_name)[1] == ".safetensors" else torch.load(file_name, map_location="cpu")
    for key in list(sd.keys()):
        if isinstance(sd[key], torch.Tensor):
            sd[key] = sd[key].to(dtype)
    return sd, {}


def merge_to_sd_model(text_encoder, unet, models, ratios, merge_dtype=torch.float32, device="cuda"):
    if merge_dtype != torch.float32:
        raise ValueError("aq_lora_merge accumulates in fp32; pass merge_dtype=torch.float and cast the result when saving")
    unet.to(merge_dtype)
    name_to_module = {}
    for name, module in unet.named_modules():
        if module.__class__.__name__ in UNET_TARGET_REPLACE_MODULE:
            for child_name, child in module.named_modules():
                if isinstance(child, (torch.nn.Linear, torch.nn.Conv2d)):
                    name_to_module[(LORA_PREFIX_UNET + "." + name + "." + child_name).replace(".", "_")] = child
    dev = torch.device(device)
    for model, ratio in zip(models, ratios):
        lora_sd, _ = load_state_dict(model, merge_dtype) if isinstance(model, (str, os.PathLike)) else (model, {})
        for key in lora_sd:
            if "lora_down" not in key:
                continue
            up_key = key.replace("lora_down", "lora_up")
            alpha_key = key[: key.index("lora_down")] + "alpha"
            module_name = ".".join(key.split(".")[:-2])
            if module_name.startswith("lora_te"):
                raise ValueError("text-encoder LoRA is not produced by AquaLoRA training and is not merged here")
            if module_name not in name_to_module:
                print(f"no module found for LoRA weight: {key}")
                continue
            module = name_to_module[module_name]
            down, up = lora_sd[key].float(), lora_sd[up_key].float()
            if down.dim() == 4 and tuple(down.shape[2:]) != (1, 1):
                raise ValueError(f"{key}: 3x3 LoRA convolutions are not AquaLoRA targets")
            dim = down.shape[0]
            alpha = float(lora_sd.get(alpha_key, dim))
            w = module.weight.data
            wd = w.float().to(dev).contiguous()
            ops.lora_merge_(wd.view(w.shape[0], -1), up.reshape(up.shape[0], dim).to(dev), down.reshape(dim, -1).to(dev), ratio * alpha / dim)
            module.weight = torch.nn.Parameter(wd.view_as(w).to(w.device, w.dtype), requires_grad=False)
    return unet


def merge(args):
    assert len(args.models) == len(args.ratios), "number of models must be equal to number of ratios"
    from aqualora_b200.unet import UNet2DConditionModel, UNetConfig

    cfg = UNetConfig.sd21() if args.v2 else UNetConfig.sd15()
    unet = UNet2DConditionModel(cfg)
    sd, _ = load_state_dict(args.sd_model, torch.float32)
    unet.load_state_dict({k[len("unet."):] if k.startswith("unet.") else k: v for k, v in sd.items()}, strict=True)
    merge_to_sd_model(None, unet, args.models, args.ratios, torch.float32)
    save_dtype = {"float": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16, None: torch.float32}[args.save_precision]
    out = {k: v.to(save_dtype).contiguous() for k, v in unet.state_dict().items()}
    save_file(out, args.save_to) if args.save_to.endswith(".safetensors") else torch.save(out, args.save_to)


def setup_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    parser.add_argument("--v2", action="store_true", help="SD 2.x U-Net topology")
    parser.add_argument("--save_precision", type=str, default=None, choices=[None, "float", "fp16", "bf16"])
    parser.add_argument("--precision", type=str, default="float", choices=["float", "fp16", "bf16"],
                        help="accepted for CLI compatibility with the reference; the merge always accumulates in fp32 (aq_lora_merge)")
    parser.add_argument("--no_metadata", action="store_true",
                        help="accepted for CLI compatibility; this script never writes sai_model_spec metadata (scripts/lib/sai_model_spec.py "
                             "stays the reference's: checkpoint I/O is out of scope)")
    parser.add_argument("--sd_model", type=str, required=True, help="diffusers-keyed U-Net state dict (.safetensors / .pt)")
    parser.add_argument("--save_to", type=str, required=True)
    parser.add_argument("--models", type=str, nargs="*", help="LoRA files (kohya / A1111 names, see diffusers_lora_to_webui.py)")
    parser.add_argument("--ratios", type=float, nargs="*")
    return parser


if __name__ == "__main__":
    merge(setup_parser().parse_args())
