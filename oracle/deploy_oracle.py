"""CPU restatement of the two "message -> deployable weights" steps (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

fold_message : scripts/create_wm_lora.py:9-51 -- down' = diag(mapper(msg)) @ down * scale (linear targets, :28-31) or
               down * m[:, None, None, None] * scale (conv targets, :33-37); `up` copied; text-encoder keys dropped (:38-39).
merge_delta  : scripts/merge_lora.py:98-120 -- W + ratio * (up @ down) * (alpha / dim) for linear and 1x1-conv modules.
Pinned by tests/golden/create_wm_lora.pt (the reference function itself, run by tools/gen_golden.py).
"""
from __future__ import annotations

import torch


def mapper_forward(emb: torch.Tensor, msg: torch.Tensor) -> torch.Tensor:
    """utils/models.py:110-115."""
    bits = emb.shape[0]
    return (emb[None] * msg[:, :, None]).sum(dim=1) / bits ** 0.5 + 1.0


def fold_message(lora_sd: dict, emb: torch.Tensor, hidinfo: str, scale: float) -> dict:
    msg = torch.tensor([int(c) for c in hidinfo]).unsqueeze(0).float()
    m = mapper_forward(emb, msg)
    out = {}
    for key, val in lora_sd.items():
        if "unet" in key:
            if "attn" in key or "ff" in key:
                if "up.weight" in key:
                    out[key] = val
                elif "down.weight" in key:
                    out[key] = torch.diag_embed(m)[0] @ val * scale
            if "proj_in" in key or "proj_out" in key:
                if "up.weight" in key:
                    out[key] = val
                elif "down.weight" in key:
                    out[key] = val * m[0][(slice(None),) + (None,) * (val.dim() - 1)] * scale
        elif "text_encoder" in key:
            pass
        else:
            raise ValueError(f"key {key} not found")
    return out


def merge_delta(weight: torch.Tensor, up: torch.Tensor, down: torch.Tensor, ratio: float, alpha=None) -> torch.Tensor:
    dim = down.shape[0]
    scale = (dim if alpha is None else alpha) / dim
    if weight.dim() == 2:
        if up.dim() == 4:
            up, down = up.squeeze(3).squeeze(2), down.squeeze(3).squeeze(2)
        return weight + ratio * (up @ down) * scale
    return weight + ratio * (up.squeeze(3).squeeze(2) @ down.squeeze(3).squeeze(2)).unsqueeze(2).unsqueeze(3) * scale
