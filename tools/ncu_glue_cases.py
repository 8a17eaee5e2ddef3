"""Launch the U-Net glue kernels on BASELINE shapes (for ncu): GroupNorm+SiLU fwd/bwd, LayerNorm fwd/bwd, GEGLU fwd/bwd.

    ncu --set full --import-source on --clock-control none -k regex:"gn_|layer_norm|geglu" -o gpurun_out/glue python tools/ncu_glue_cases.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aqualora_b200 import unet_ops

GN = [(16, 320, 64, 64), (16, 640, 32, 32), (16, 1280, 16, 16), (16, 1920, 32, 32), (16, 1280, 8, 8)]
LN = [(16 * 4096, 320), (16 * 1024, 640), (16 * 256, 1280)]
GG = [(16 * 4096, 1280), (16 * 256, 5120)]


def main():
    dev = torch.device("cuda:0")
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    g = torch.Generator(device=dev).manual_seed(0)
    for (B, C, H, W) in GN:
        x = torch.randn(B, C, H, W, generator=g, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        gam = torch.ones(C, device=dev, dtype=torch.bfloat16)
        bet = torch.zeros(C, device=dev, dtype=torch.bfloat16)
        t = torch.randn(B, C, generator=g, device=dev).bfloat16()
        dy = torch.randn(B, C, H, W, generator=g, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        for _ in range(reps):
            y = unet_ops.group_norm_nhwc(x, gam, bet, 32, 1e-5, True, t)
            y.backward(dy)
            x.grad = None
    for (M, C) in LN:
        x = torch.randn(M, C, generator=g, device=dev).bfloat16().requires_grad_(True)
        gam = torch.ones(C, device=dev, dtype=torch.bfloat16)
        bet = torch.zeros(C, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(M, C, generator=g, device=dev).bfloat16()
        for _ in range(reps):
            y = unet_ops.layer_norm(x, gam, bet, 1e-5)
            y.backward(dy)
            x.grad = None
    for (M, F) in GG:
        p = torch.randn(M, 2 * F, generator=g, device=dev).bfloat16().requires_grad_(True)
        go = torch.randn(M, F, generator=g, device=dev).bfloat16()
        for _ in range(reps):
            y = unet_ops.geglu(p)
            y.backward(go)
            p.grad = None
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
