"""Launch the fused projection+LoRA kernel on a few BASELINE shapes (for `ncu --set full -k regex:lora_gemm`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aqualora_b200 import ops

SHAPES = [(65536, 320, 320, 4096), (4096, 1280, 1280, 256), (16384, 640, 5120, 1024), (65536, 1280, 320, 4096)]


def main():
    dev = torch.device("cuda:0")
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    shapes = SHAPES
    if len(sys.argv) > 2:      # "M,K,N,tok;M,K,N,tok"
        shapes = [tuple(int(v) for v in item.split(",")) for item in sys.argv[2].split(";")]
    for (M, K, N, tok) in shapes:
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(M, K, generator=g, device=dev).bfloat16()
        w = (torch.randn(N, K, generator=g, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(N, generator=g, device=dev).bfloat16()
        dn = (torch.randn(64, K, generator=g, device=dev) * K ** -0.5).bfloat16()
        up = (torch.randn(N, 64, generator=g, device=dev) * 0.1).bfloat16()
        sc = torch.randn(M // tok, 64, generator=g, device=dev)
        for _ in range(reps):
            ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
