"""Bit accuracy / TPR of decoded messages: the metric definition of the reference's `evaluation/utils_eval.py`
(`calculate_fpr`, `get_threshold` :131-140; `simple_decode` :156-213), on the CUDA decoder and batched.

The reference decodes one image per forward (batch 1, one host->device copy each); here a batch of images goes through
`SecretDecoder.decode_bits` (csrc/decoder.cu) at once.  The statistics are integer arithmetic and are identical by construction:
acc_i = matches_i / bits, bit accuracy = mean_i acc_i, TPR = #{acc_i >= tau / bits} / n with tau the smallest threshold whose
binomial false-positive rate is <= the requested FPR.
"""
from __future__ import annotations

from math import comb
from typing import List, Optional, Sequence, Tuple, Union

import torch


def calculate_fpr(tau: int, k: int) -> float:
    """P[more than tau of k fair coin flips match] -- evaluation/utils_eval.py:131-134."""
    return sum(comb(k, i) for i in range(tau + 1, k + 1)) / (2 ** k)


def get_threshold(k: int, fpr: float) -> int:
    """Smallest tau with calculate_fpr(tau, k) <= fpr -- evaluation/utils_eval.py:136-140 (48 bits: 35 at 1e-3, 40 at 1e-6)."""
    tau = 0
    while calculate_fpr(tau, k) > fpr:
        tau += 1
    return tau


def _as_bits(msg: Union[str, Sequence[int], torch.Tensor], bits: int, device) -> torch.Tensor:
    if isinstance(msg, str):
        msg = [int(c) for c in msg]
    t = torch.as_tensor(msg, device=device).reshape(-1, bits) if not isinstance(msg, torch.Tensor) else msg.to(device).reshape(-1, bits)
    return t.to(torch.uint8)


def bit_accuracy(pred_bits: torch.Tensor, msg_gt, tpr_threshold: float = 1e-3) -> Tuple[float, float, torch.Tensor]:
    """(bit accuracy, TPR, per-image accuracy) of decoded bits [n, k] against one ground-truth message (string / list / [k]) or one
    message per image ([n, k]) -- evaluation/utils_eval.py:197-211."""
    n, k = pred_bits.shape
    gt = _as_bits(msg_gt, k, pred_bits.device)
    if gt.shape[0] not in (1, n):
        raise ValueError(f"ground truth holds {gt.shape[0]} messages for {n} images")
    matches = (pred_bits.to(torch.uint8) == gt).sum(dim=1)            # integer counts: exact
    acc = matches.to(torch.float64) / k
    tau = get_threshold(k, tpr_threshold)
    tp = int((matches >= tau).sum())                                   # acc >= tau / k  <=>  matches >= tau
    return float(acc.mean()), tp / n, acc


def _load_images(items, resolution: int) -> torch.Tensor:
    """PIL images / paths -> [n, 3, res, res] fp32 in [-1, 1] exactly as `process` of evaluation/utils_eval.py:171-178 (RGB, bicubic
    resize on the host, / 127.5 - 1)."""
    import numpy as np
    from PIL import Image

    out = []
    for im in items:
        if not isinstance(im, Image.Image):
            im = Image.open(im)
        if im.mode != "RGB":
            im = im.convert("RGB")
        im = im.resize((resolution, resolution), resample=Image.Resampling.BICUBIC)
        a = (np.array(im).astype(np.uint8) / 127.5 - 1.0).astype(np.float32)
        out.append(torch.from_numpy(a).permute(2, 0, 1))
    return torch.stack(out)


def simple_decode(bitnum: int, msgdecoder, images, msg_gt=None, resolution: int = 512, tpr_threshold: float = 1e-3,
                  batch_size: int = 64, device="cuda") -> Tuple[Optional[float], Optional[float], List[str]]:
    """`simple_decode` of evaluation/utils_eval.py:156-213.  `msgdecoder`: a path to `msgdecoder.pt` or a `SecretDecoder`;
    `images`: a list of PIL images / paths (the reference's input) or a [n, 3, H, W] tensor in [-1, 1].
    Returns (bit accuracy, TPR, decoded bit strings); the first two are None without `msg_gt`."""
    from .decoder import SecretDecoder

    dev = torch.device(device)
    if not isinstance(msgdecoder, torch.nn.Module):
        dec = SecretDecoder(output_size=bitnum)
        dec.load_state_dict(torch.load(msgdecoder, map_location="cpu"))
        msgdecoder = dec
    msgdecoder = msgdecoder.to(dev).eval()
    x = images if isinstance(images, torch.Tensor) else _load_images(images, resolution)
    bits = []
    for i in range(0, x.shape[0], batch_size):
        bits.append(msgdecoder.decode_bits(x[i:i + batch_size].to(dev, non_blocking=True)))
    bits = torch.cat(bits).view(-1, bitnum)
    results = ["".join(map(str, row)) for row in bits.cpu().tolist()]
    if msg_gt is None:
        return None, None, results
    bitacc, tpr, _ = bit_accuracy(bits, msg_gt, tpr_threshold)
    return bitacc, tpr, results
