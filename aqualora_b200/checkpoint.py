"""Checkpoint artefacts of the PPFT stage, in the reference's file formats (SURVEY.md 8(b) "state-dict / file contract").

    <dir>/pytorch_lora_weights.safetensors   {'unet.<renamed key>.{down,up}.weight': fp32}    train/ppft_train.py:443-471, :699-725, :1203-1220
    <dir>/mapper.pt                          {'bit_embeddings.weight': [bits, r]}             :716-717, :1226-1228
    <dir>/msgdecoder.pt                      SecretDecoder.state_dict() (keys under `model.`)  :714-715, :1222-1224
    <dir>/trainer_state.pt                   flat AdamW moments + step counter (what `accelerator.save_state` keeps in
                                             optimizer.bin / scheduler.bin so a run can resume: :943-965, :1079-1103)

`scripts/create_wm_lora.py` and `scripts/merge_lora.py` consume the first two unchanged.  Host-side plumbing only (no kernels);
works on CPU tensors as well so the format is testable without a GPU.
"""
from __future__ import annotations

import os
import re
import shutil
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

from .lora_modules import lora_state_dict_key, resolve, unet_attn_processors_state_dict

LORA_WEIGHT_NAME = "pytorch_lora_weights.safetensors"     # diffusers' LORA_WEIGHT_NAME_SAFE


def lora_state_dict_for_save(unet: nn.Module, keys: Iterable[str]) -> Dict[str, torch.Tensor]:
    """`LoraLoaderMixin.save_lora_weights(unet_lora_layers=unet_attn_processors_state_dict(unet))`: every entry prefixed `unet.`
    (the prefix scripts/create_wm_lora.py:25 filters on), fp32, contiguous, on the CPU."""
    return {f"unet.{k}": v.detach().to(torch.float32).cpu().contiguous().clone()
            for k, v in unet_attn_processors_state_dict(unet, keys).items()}


def save_lora_weights(save_directory: str, unet: nn.Module, keys: Iterable[str]) -> str:
    from safetensors.torch import save_file

    os.makedirs(save_directory, exist_ok=True)
    path = os.path.join(save_directory, LORA_WEIGHT_NAME)
    save_file(lora_state_dict_for_save(unet, keys), path, metadata={"format": "pt"})
    return path


def load_lora_into_unet(state_dict_or_dir, unet: nn.Module, keys: Iterable[str], strict: bool = True) -> int:
    """Inverse of `save_lora_weights` (`load_model_hook`, train/ppft_train.py:727-745; also the `--resume_from_lora` warm start,
    :626-633,670-671): copies `unet.<renamed key>.{down,up}.weight` into the matching `lora_layer` parameters IN PLACE (so views
    into a flat parameter buffer stay views).  Returns the number of tensors loaded."""
    if isinstance(state_dict_or_dir, (str, os.PathLike)):
        from safetensors.torch import load_file

        p = os.fspath(state_dict_or_dir)
        sd = load_file(os.path.join(p, LORA_WEIGHT_NAME) if os.path.isdir(p) else p, device="cpu")
    else:
        sd = state_dict_or_dir
    n = 0
    used = set()
    for key in keys:
        lora = resolve(unet, key).lora_layer
        stem = "unet." + lora_state_dict_key(key)
        for which in ("down", "up"):
            name = f"{stem}.{which}.weight"
            if name not in sd:
                if strict:
                    raise KeyError(f"{name} is missing from the LoRA checkpoint")
                continue
            dst = getattr(lora, which).weight
            src = sd[name]
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{name}: checkpoint shape {tuple(src.shape)} != module shape {tuple(dst.shape)}")
            with torch.no_grad():
                dst.copy_(src.to(device=dst.device, dtype=dst.dtype))
            used.add(name)
            n += 1
    if strict:
        extra = [k for k in sd if k.startswith("unet.") and k not in used]
        if extra:
            raise KeyError(f"{len(extra)} unexpected unet.* entries in the LoRA checkpoint, e.g. {extra[0]}")
    return n


def save_ppft_artifacts(output_dir: str, unet: nn.Module, keys: Iterable[str], mapper_emb: torch.Tensor,
                        msgdecoder: Optional[nn.Module] = None) -> None:
    """The three files `save_model_hook` / the final save of train/ppft_train.py write."""
    save_lora_weights(output_dir, unet, keys)
    torch.save({"bit_embeddings.weight": mapper_emb.detach().to(torch.float32).cpu().clone()}, os.path.join(output_dir, "mapper.pt"))
    if msgdecoder is not None:
        torch.save({k: v.detach().cpu().clone() for k, v in msgdecoder.state_dict().items()}, os.path.join(output_dir, "msgdecoder.pt"))


def rotate_checkpoints(output_dir: str, total_limit: Optional[int]) -> None:
    """`--checkpoints_total_limit` (train/ppft_train.py:1083-1099): before saving a new `checkpoint-<step>`, delete the oldest so that
    at most `total_limit - 1` remain."""
    if total_limit is None:
        return
    cks = sorted((d for d in os.listdir(output_dir) if re.fullmatch(r"checkpoint-\d+", d)), key=lambda d: int(d.split("-")[1]))
    if len(cks) >= total_limit:
        for d in cks[:len(cks) - total_limit + 1]:
            shutil.rmtree(os.path.join(output_dir, d))


def latest_checkpoint(output_dir: str) -> Optional[str]:
    """`--resume_from_checkpoint latest` (train/ppft_train.py:947-952)."""
    if not os.path.isdir(output_dir):
        return None
    cks = sorted((d for d in os.listdir(output_dir) if re.fullmatch(r"checkpoint-\d+", d)), key=lambda d: int(d.split("-")[1]))
    return os.path.join(output_dir, cks[-1]) if cks else None
