"""Where does a PPFT step's device time go, by ATen op (with shapes and the issuing Python frame) and by kernel name?"""
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
    ka = prof.key_averages(group_by_stack_n=8, group_by_input_shape=True)
    rows = [e for e in ka if e.self_device_time_total > 0 and e.key.startswith("aten::")]
    rows.sort(key=lambda e: -e.self_device_time_total)
    print("== ATen ops by self device time (top 45)")
    for e in rows[:45]:
        print(f"{e.key:28s} {e.self_device_time_total / 1e3:8.3f} ms  x{e.count:4d}  shapes {str(e.input_shapes)[:100]}")
        shown = 0
        for fr in e.stack:
            if "aqualora_b200" in fr and shown < 2:
                print("      ", fr[:170])
                shown += 1
    kern = {}
    for e in prof.events():
        if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower():
            k = kern.setdefault(e.name[:90], [0.0, 0])
            k[0] += e.device_time; k[1] += 1
    tot = sum(v[0] for v in kern.values())
    print(f"== kernels by device time (total {tot / 1e3:.2f} ms)")
    for name, (t, n) in sorted(kern.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f} %  x{n:4d}  {name}")


if __name__ == "__main__":
    main()
