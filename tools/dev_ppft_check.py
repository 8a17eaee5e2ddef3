"""Developer harness: PPFT step on a B200 -- gradient parity of the tiny U-Net against the CPU oracle, then timing and a
kernel-time breakdown of the SD1.5 B=16 step (and the PyTorch-eager GPU sequence for context)."""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def synth_batch(B, cfg, device, seed, bits=48, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    s = cfg.sample_size
    lat = (torch.randn(B, 4, s, s, generator=g) * 0.18215)
    wm = torch.randn(B, 4, s, s, generator=g) * 0.02 * 0.18215
    noise = torch.randn(B, 4, s, s, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g)
    msg = torch.randint(0, 2, (B, bits), generator=g).float()
    f = lambda x: x.to(device=device, dtype=dtype)
    return f(lat), f(wm), f(noise), t.to(device), f(ctx), msg.to(device)


def parity_tiny():
    from aqualora_b200 import ppft
    from aqualora_b200.unet import UNetConfig
    from oracle import lora_oracle as O
    from oracle.patch import patch_with_oracle

    dev = torch.device("cuda:0")
    cfg = UNetConfig.tiny(16)
    rank, bits, B = 8, 48, 2
    unet = ppft.build_unet(cfg, dev, seed=3)
    emb = O.mapper_init(bits, rank, generator=torch.Generator().manual_seed(5))
    tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=rank, msg_bits=bits), emb, dev, lora_up_std=0.05, seed=1)
    batch = synth_batch(B, cfg, dev, 11)
    loss = tr.forward_backward(*batch)
    torch.cuda.synchronize()
    g_flat = tr.state.grad.clone()

    # oracle: same module tree on CPU in fp32 (weights = the bf16 values), reference op sequence via monkey patch
    cpu_unet = ppft.build_unet(cfg, "cpu", dtype=torch.float32, seed=3)
    from aqualora_b200 import lora_modules
    layers = lora_modules.inject_lora(cpu_unet, lora_modules_keys(cpu_unet), rank)
    sd = {k: v.detach().float().cpu() for k, v in unet.state_dict().items()}
    cpu_unet.load_state_dict(sd)
    patch_with_oracle(cpu_unet)
    for _, _, l in layers:
        l.down.weight.requires_grad_(True)
        l.up.weight.requires_grad_(True)
    E = tr.state.mapper_emb.detach().cpu().clone().requires_grad_(True)
    lat, wm, noise, t, ctx, msg = [x.detach().cpu() for x in batch]
    lat, wm, noise, ctx = lat.float(), wm.float(), noise.float(), ctx.float()
    scale = O.mapper_forward(msg, E).to(torch.bfloat16).float()
    ac = ppft.scaled_linear_alphas_cumprod()
    noisy = ppft.add_noise(ac, lat, noise, t)
    noisy_wm = ppft.add_noise(ac, lat + wm, noise, t)
    with torch.no_grad():
        clean = cpu_unet(noisy, t, ctx, cross_attention_kwargs={"scale": torch.zeros_like(scale)}).sample
    pred = cpu_unet(noisy_wm, t, ctx, cross_attention_kwargs={"scale": scale}).sample
    loss_ref = torch.nn.functional.mse_loss(pred, clean)
    loss_ref.backward()
    g_ref = torch.cat([p.grad.reshape(-1) for _, _, l in layers for p in (l.down.weight, l.up.weight)])
    g_got = g_flat[:g_ref.numel()].cpu()
    cos = torch.nn.functional.cosine_similarity(g_got, g_ref, dim=0).item()
    rel = ((g_got - g_ref).norm() / g_ref.norm()).item()
    ge_got = tr.state.mapper_grad.cpu().reshape(-1)
    ge_ref = E.grad.reshape(-1)
    res = {"case": "parity_tiny", "loss": loss.item(), "loss_ref": loss_ref.item(), "grad_cos": cos, "grad_rel_err": rel,
           "mapper_grad_cos": torch.nn.functional.cosine_similarity(ge_got, ge_ref, dim=0).item(),
           "mapper_grad_rel": ((ge_got - ge_ref).norm() / ge_ref.norm()).item(),
           "gnorm": g_ref.norm().item()}
    # optimizer step parity (fp32 AdamW on the oracle grads vs the fused kernel on ours)
    p_before = tr.state.param.clone()
    tr.optimizer_step()
    torch.cuda.synchronize()
    params = [p for _, _, l in layers for p in (l.down.weight, l.up.weight)]
    opt = torch.optim.AdamW([{"params": params}, {"params": [E]}], lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    with torch.no_grad():
        off = 0
        for p in params:  # same starting point and same grads as the device run -> isolates the update arithmetic
            p.copy_(p_before[off:off + p.numel()].view_as(p).cpu())
            p.grad.copy_(g_flat[off:off + p.numel()].view_as(p).cpu())
            off += p.numel()
        E.grad.copy_(g_flat[tr.state.mapper_off:tr.state.mapper_off + E.numel()].view_as(E).cpu())
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()
    p_ref = torch.cat([p.detach().reshape(-1) for p in params])
    p_got = tr.state.param[:p_ref.numel()].cpu()
    res["adamw_max_abs_diff"] = (p_got - p_ref).abs().max().item()
    res["adamw_mapper_max_abs_diff"] = (tr.state.mapper_emb.cpu() - E.detach()).abs().max().item()
    res["grad_zeroed"] = bool((tr.state.grad == 0).all().item())
    res["ok"] = cos > 0.99 and rel < 0.1 and res["adamw_max_abs_diff"] < 1e-6
    print("RESULT " + json.dumps(res))


def lora_modules_keys(unet):
    from aqualora_b200.unet import lora_target_keys
    return lora_target_keys(unet)


def time_sd15(B, steps, profile):
    from aqualora_b200 import ppft
    from aqualora_b200.unet import UNetConfig
    from oracle import lora_oracle as O

    dev = torch.device("cuda:0")
    cfg = UNetConfig.sd15(64)
    t0 = time.time()
    unet = ppft.build_unet(cfg, dev, seed=0)
    emb = O.mapper_init(48, 64, generator=torch.Generator().manual_seed(5))
    tr = ppft.PPFTTrainer(unet, ppft.PPFTConfig(rank=64), emb, dev, lora_up_std=0.02, seed=1)
    print(f"build {time.time() - t0:.1f}s, lora params {tr.state.n_lora}", flush=True)
    batch = synth_batch(B, cfg, dev, 1234)
    for _ in range(3):
        loss = tr.step(*batch)
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = tr.step(*batch)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    res = {"case": f"sd15_B{B}", "ms_per_step": med, "img_per_s": B / med * 1e3, "loss": loss.item(),
           "mem_GB": torch.cuda.max_memory_allocated() / 1e9, "all_ms": [round(t, 1) for t in ts]}
    print("RESULT " + json.dumps(res), flush=True)
    if profile:
        from torch.profiler import ProfilerActivity, profile as tprof
        with tprof(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            tr.step(*batch)
            torch.cuda.synchronize()
        tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=100, max_name_column_width=120)
        print(tab)
    return tr, batch


def time_eager_gpu(B, steps):
    """The reference's unfused op sequence on the same GPU (PyTorch eager: cuBLAS + bmm + add), bf16 autocast-like."""
    from aqualora_b200 import lora_modules, ppft
    from aqualora_b200.unet import UNetConfig, lora_target_keys
    from oracle import lora_oracle as O
    from oracle.patch import patch_with_oracle

    dev = torch.device("cuda:0")
    cfg = UNetConfig.sd15(64)
    unet = ppft.build_unet(cfg, dev, seed=0)
    layers = lora_modules.inject_lora(unet, lora_target_keys(unet), 64)
    g = torch.Generator().manual_seed(1)
    params = []
    for _, _, l in layers:
        l.up.weight.data.copy_(torch.randn(l.up.weight.shape, generator=g) * 0.02)
        l.to(torch.bfloat16)   # eager arm: bf16 LoRA weights (what autocast would feed cuBLAS)
        for p in (l.down.weight, l.up.weight):
            p.requires_grad_(True)
            params.append(p)
    patch_with_oracle(unet)
    emb = O.mapper_init(48, 64, generator=torch.Generator().manual_seed(5)).to(dev).requires_grad_(True)
    opt = torch.optim.AdamW(params + [emb], lr=1e-4, weight_decay=1e-2)
    lat, wm, noise, t, ctx, msg = synth_batch(B, cfg, dev, 1234)
    ac = ppft.scaled_linear_alphas_cumprod().to(dev)

    def step():
        scale = O.mapper_forward(msg, emb).to(torch.bfloat16)
        noisy = ppft.add_noise(ac, lat, noise, t)
        noisy_wm = ppft.add_noise(ac, lat + wm, noise, t)
        clean = unet(noisy, t, ctx, cross_attention_kwargs={"scale": torch.zeros_like(scale)}).sample.detach()
        pred = unet(noisy_wm, t, ctx, cross_attention_kwargs={"scale": scale}).sample
        loss = torch.nn.functional.mse_loss(pred.float(), clean.float())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print("RESULT " + json.dumps({"case": f"eager_gpu_sd15_B{B}", "ms_per_step": med, "img_per_s": B / med * 1e3,
                                  "loss": loss.item(), "mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="parity,sd15")
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    if "parity" in a.what:
        parity_tiny()
    if "sd15" in a.what:
        time_sd15(a.B, a.steps, a.profile)
    if "eager" in a.what:
        time_eager_gpu(a.B, a.steps)
