"""Run the decoder chain once or twice on a batch (for `ncu -k regex:'pointwise|depthwise|stem|se_kernel|fc_kernel'`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aqualora_b200.decoder import SecretDecoder


def main():
    dev = torch.device("cuda:0")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.manual_seed(0)
    dec = SecretDecoder(48).eval()      # parameters stay on the host: packing (BN fold) then runs on the CPU, only our kernels hit the GPU
    x = torch.rand(B, 3, 512, 512, device=dev) * 2 - 1
    for _ in range(reps):
        dec.decode_bits(x)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
