"""Drop-in `scripts/create_wm_lora.py` (reference file of the same name): fold one message into a trained watermark LoRA.

Same entry point and CLI -- `create_watermark_lora(train_folder, scale, msg_bits=48, hidinfo=None, save=True)`,
`--train_folder --msg_bits --scale --hidinfo` -- same input / output files (`pytorch_lora_weights.safetensors`, `mapper.pt`,
`<train_folder>/<bits>/pytorch_lora_weights.safetensors`).  The arithmetic runs on the B200: `aq_mapper_fwd` for
m = mapper(msg) and `aq_lora_fold_down` for down' = (down * m[:, None]) * scale (csrc/lora_deploy.cu); there is no CPU path.
Two reference limitations are lifted: the LoRA rank is read from `mapper.pt` instead of being hard-coded to 320
(create_wm_lora.py:19), and `proj_in/proj_out` may be linear (SD 2.x, 2-D weights) as well as 1x1 convolutions (:32-37).
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from safetensors.torch import load_file, save_file  # noqa: E402

from aqualora_b200 import ops  # noqa: E402


def create_watermark_lora(train_folder, scale, msg_bits=48, hidinfo=None, save=True, device="cuda"):
    lora_state_dict = load_file(f"{train_folder}/pytorch_lora_weights.safetensors", device="cpu")
    if hidinfo is None:
        hidinfo = torch.randint(0, 2, (1, msg_bits))
    else:
        assert len(hidinfo) == msg_bits
        hidinfo = torch.tensor([int(i) for i in hidinfo]).unsqueeze(0)
    dev = torch.device(device)
    emb = torch.load(f"{train_folder}/mapper.pt", map_location="cpu")["bit_embeddings.weight"].float().to(dev)   # [bits, r]
    assert emb.shape[0] == msg_bits, f"mapper.pt holds {emb.shape[0]}-bit embeddings, asked for {msg_bits}"
    m = ops.mapper_fwd(hidinfo.float().to(dev), emb.contiguous(), round_bf16=False)[0].contiguous()              # [r], fp32

    c_lora_state_dict = {}
    for key, val in lora_state_dict.items():
        if "unet" in key:
            if not any(t in key for t in ("attn", "ff", "proj_in", "proj_out")):
                continue                                              # the reference silently drops such keys too
            if "up.weight" in key:
                c_lora_state_dict[key] = val
            elif "down.weight" in key:
                folded = ops.lora_fold_down(val.float().to(dev), m, scale)
                c_lora_state_dict[key] = folded.to(val.dtype).cpu()
        elif "text_encoder" in key:
            pass
        else:
            raise ValueError(f"key {key} not found")

    hidinfo = "".join(map(str, hidinfo.tolist()[0]))
    if save:
        os.makedirs(f"{train_folder}/{hidinfo}", exist_ok=True)
        save_file(c_lora_state_dict, f"{train_folder}/{hidinfo}/pytorch_lora_weights.safetensors")
    return hidinfo, c_lora_state_dict


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--train_folder", type=str, required=True)
    parser.add_argument("--msg_bits", type=int, default=48)
    parser.add_argument("--scale", type=float, default=1.03)
    parser.add_argument("--hidinfo", type=str, default=None, help="your secret message, if None, it will be randomly generated")
    args = parser.parse_args()
    hidinfo, _ = create_watermark_lora(args.train_folder, args.scale, args.msg_bits, args.hidinfo)
    print(hidinfo)
