"""One eager PPFT step (SD1.5, B = 16) between cudaProfilerStart / Stop, for an ncu launch list of the whole step:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/step_launches.csv python tools/ncu_step.py
    python tools/launch_list.py gpurun_out/step_launches.csv --aggregate-only
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools.dev_ppft_check import synth_batch


def main():
    from aqualora_b200 import ppft
    from aqualora_b200.unet import UNetConfig
    from oracle import lora_oracle as O

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda:0")
    cfg = UNetConfig.sd15(64)
    unet = ppft.build_unet(cfg, dev, seed=0)
    emb = O.mapper_init(48, 64, generator=torch.Generator().manual_seed(5))
    tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=64), emb, dev, lora_up_std=0.02, seed=1)
    batch = synth_batch(B, cfg, dev, 1234)
    for _ in range(3):
        tr.step(*batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.step(*batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
