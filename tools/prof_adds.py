"""Which Python lines issue the elementwise add / copy kernels of a PPFT step (torch profiler with stacks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from tools.dev_ppft_check import synth_batch


def main():
    from aqualora_b200 import ppft
    from aqualora_b200.unet import UNetConfig
    from oracle import lora_oracle as O

    dev = torch.device("cuda:0")
    cfg = UNetConfig.sd15(64)
    unet = ppft.build_unet(cfg, dev, seed=0)
    emb = O.mapper_init(48, 64, generator=torch.Generator().manual_seed(5))
    tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=64), emb, dev, lora_up_std=0.02, seed=1)
    batch = synth_batch(16, cfg, dev, 1234)
    for _ in range(3):
        tr.step(*batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
        tr.step(*batch)
        torch.cuda.synchronize()
    ka = prof.key_averages(group_by_stack_n=6, group_by_input_shape=True)
    rows = [e for e in ka if e.key in ("aten::add", "aten::add_", "aten::copy_", "aten::mul", "aten::cat", "aten::sum", "aten::div", "aten::to",
                                       "aten::_to_copy", "aten::upsample_nearest2d", "aten::silu", "aten::fill_", "aten::zero_")]
    rows.sort(key=lambda e: -e.self_device_time_total)
    for e in rows[:40]:
        print(f"{e.key:18s} cuda {e.self_device_time_total / 1e3:8.3f} ms  x{e.count:4d}  shapes {str(e.input_shapes)[:90]}")
        for fr in e.stack[:6]:
            if "aqualora_b200" in fr or "autograd" in fr:
                print("      ", fr[:160])


if __name__ == "__main__":
    main()
