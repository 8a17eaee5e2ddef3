"""1-vs-N GPU equivalence of the PPFT gradient exchange (SURVEY.md 4 "DDP": same global batch, same seeds -> same LoRA gradients).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_grad_check.py --out f.json

Every rank builds the same U-Net + LoRA state, takes its slice of ONE global batch, runs forward + backward and the single flat
all-reduce (PPFTTrainer.exchange_gradients); rank 0 then recomputes the whole batch alone and compares the flat gradients."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--model", default="tiny", choices=["tiny", "sd15"])
    ap.add_argument("--per-rank", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    from aqualora_b200 import ppft
    from aqualora_b200.unet import UNetConfig

    cfg = UNetConfig.tiny(16) if a.model == "tiny" else UNetConfig.sd15(64)
    r = 8 if a.model == "tiny" else 64

    def trainer():
        unet = ppft.build_unet(cfg, dev, seed=3)
        emb = torch.randn(48, r, generator=torch.Generator().manual_seed(5))
        return ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=r), emb, dev, lora_up_std=0.05, seed=1)

    B = a.per_rank * world
    g = torch.Generator().manual_seed(11)
    s = cfg.sample_size
    lat = torch.randn(B, 4, s, s, generator=g) * 0.18215
    wm = torch.randn(B, 4, s, s, generator=g) * 0.02
    noise = torch.randn(B, 4, s, s, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g)
    msg = torch.randint(0, 2, (B, 48), generator=g).float()
    bf = lambda x: x.to(dev, torch.bfloat16)
    sl = slice(rank * a.per_rank, (rank + 1) * a.per_rank)
    tr = trainer()
    loss = tr.forward_backward(bf(lat[sl]), bf(wm[sl]), bf(noise[sl]), t[sl].to(dev), bf(ctx[sl]), msg[sl].to(dev))
    scale = tr.exchange_gradients()                       # ONE all-reduce of the flat LoRA (+ mapper) gradient buffer
    g_dist = (tr.state.grad * scale).clone()
    loss_all = loss.clone()
    dist.all_reduce(loss_all)
    res = None
    if rank == 0:
        one = trainer()
        one.world = 1
        loss1 = one.forward_backward(bf(lat), bf(wm), bf(noise), t.to(dev), bf(ctx), msg.to(dev))
        g_one = one.state.grad
        cos = torch.nn.functional.cosine_similarity(g_dist, g_one, dim=0).item()
        rel = ((g_dist - g_one).norm() / g_one.norm()).item()
        res = {"world": world, "model": a.model, "global_batch": B, "grad_cosine": cos, "grad_rel_fro": rel, "grad_norm": g_one.norm().item(),
               "loss_single": loss1.item(), "loss_mean_over_ranks": (loss_all / world).item(), "flat_grad_floats": g_one.numel()}
        print(json.dumps(res), flush=True)
        if a.out:
            json.dump(res, open(a.out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
