"""One GPU pretrain step (train/latent_wm_pretrain.py:164-217, B = 2, stub VAE / LPIPS) between cudaProfilerStart / Stop, for an ncu
launch list: which kernels the training side of hot path (ii) runs on.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/pretrain_launches.csv python tools/ncu_pretrain.py
"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    from aqualora_b200 import noise_layers, pretrain
    from aqualora_b200.decoder import SecretDecoder
    from aqualora_b200.models import SecretEncoder
    from oracle.pretrain_oracle import StubVAE, lpips_stub          # frozen third-party stand-ins (plain torch)

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = SecretEncoder(48).to(dev).train()
    torch.nn.init.normal_(enc.secret_scaler[5].weight, std=0.05)
    dec = SecretDecoder(48).to(dev).train()
    vae = StubVAE(0).to(dev)
    noiser = noise_layers.Noiser(["Jpeg", "CropandResize", "GaussianBlur", "GaussianNoise", "ColorJitter"], [0.4, 0.1, 0.2, 0.05, 0.1, 0.15], dev,
                                 rng=np.random.default_rng(7))
    g = torch.Generator().manual_seed(1)
    image = (torch.rand(B, 3, 512, 512, generator=g) * 2 - 1).to(dev)
    msg = torch.randint(0, 2, (B, 48), generator=g).to(dev)
    opt = torch.optim.AdamW(list(enc.parameters()) + list(dec.parameters()), lr=1e-3, weight_decay=1e-4)
    rng = random.Random(3)

    def step():
        opt.zero_grad()
        out = pretrain.pretrain_step(enc, dec, vae.encode, vae.decode, lpips_stub, noiser, image, msg, rng, [0.4, 0.1, 0.2, 0.05, 0.1, 0.15], False, 2)
        opt.step()
        return out

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"pretrain step B={B}: {e0.elapsed_time(e1) / 5:.2f} ms / step (eager issue)")
    torch.cuda.cudart().cudaProfilerStart()
    out = step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print({k: float(v) for k, v in out.items() if k in ("loss", "msgloss", "lpips", "prvl")})


if __name__ == "__main__":
    main()
