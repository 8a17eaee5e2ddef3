"""Minimal 2+-rank NCCL check: init, one all_reduce, one barrier, timings (diagnoses environment hangs apart from bench.py)."""
import os
import time

import torch
import torch.distributed as dist


def main():
    t0 = time.time()
    rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    print(f"[{rank}] init {time.time() - t0:.1f}s", flush=True)
    x = torch.ones(1 << 20, device=dev)
    dist.all_reduce(x)
    torch.cuda.synchronize()
    print(f"[{rank}] all_reduce ok {x[0].item()} {time.time() - t0:.1f}s", flush=True)
    big = torch.ones(27_200_000, device=dev)
    for _ in range(3):
        dist.all_reduce(big)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(big)
    e1.record()
    torch.cuda.synchronize()
    print(f"[{rank}] 108.8 MB all_reduce: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    print(f"[{rank}] done {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
