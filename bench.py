#!/usr/bin/env python
"""Headline benchmark: PPFT images/sec, SD1.5 512x512 (64x64 latents), LoRA rank 64, 48-bit messages, bf16, per-GPU batch 16
(BASELINE.json configs[1]; configs[2] is the same step at N = 2/4/8 with one NCCL allreduce of the flat LoRA gradients).

    python bench.py [--gpus N] [--steps K] [--warmup W]                    # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]              # the reference op sequence on the host CPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic batch: mapper -> clean U-Net forward (LoRA skipped) ->
watermarked U-Net forward (192 fused projection+LoRA kernels) -> MSE -> backward (fused dX / weight-grad kernels) ->
[allreduce] -> clip + AdamW over the flat buffers (train/ppft_train.py:987-1068).  VAE / text encoder are outside the
north-star path and have no weights offline: latents and text context are synthetic.

Prints ONE JSON line (rank 0).  `value` = images/s with the batch already resident in HBM; `e2e` = the same step through the
public API with the batch in pinned host memory (H2D copies and the D2H read of the loss inside the timed region).
`roofline` describes the dominant kernel (aq::lora_gemm_kernel, the fused base-GEMM + LoRA contraction): algorithmic FLOPs
per launch / average launch duration, measured live and device-only -- the forward launches of one recorded step re-issued
inside a CUDA graph, CUDA events around its replays -- against MEASURED_PEAKS.json, with the same kernels' CUPTI durations inside
the training step beside it.  `cpu_baseline` = the oracle's PyTorch-eager restatement of the reference step on the host cores
(bounded sample: B=1); `gpu_eager_baseline` = the reference's unfused op sequence on PyTorch eager (cuBLAS / cuDNN) on the same
B200 at the same batch.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "ppft_images_per_sec"
UNIT = "images/s"
RANK_R = 64
BITS = 48
PER_GPU_BATCH = 16


# --------------------------------------------------------------------------------------------------------------------
# synthetic PPFT batch: generator seed = 1234 + step (+ rank)
# --------------------------------------------------------------------------------------------------------------------
def synth_batch(B, cfg, seed, encoder_state=None):
    g = torch.Generator().manual_seed(seed)
    s = cfg.sample_size
    lat = torch.randn(B, 4, s, s, generator=g) * 0.18215
    noise = torch.randn(B, 4, s, s, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g)
    msg = torch.randint(0, 2, (B, BITS), generator=g).float()
    return lat, noise, t, ctx, msg


def workload_name(model: str, sample_size: int) -> str:
    return (f"{'SD1.5' if model == 'sd15' else 'SD2.1-base'} PPFT step, LoRA rank {RANK_R} on the 192 unet_keys.json targets, "
            f"{BITS}-bit messages, {sample_size * 8}x{sample_size * 8} ({sample_size}x{sample_size} latents), random-init U-Net, "
            "VAE/text-encoder excluded")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------------
# roofline instrumentation: device-only.  One eagerly issued step is RECORDED (which launches, with which tensors), then each
# kind of launch is re-issued back to back inside its own CUDA graph and the replays are timed with CUDA events: nothing but
# this library's kernels runs between the two events, so a slow host cannot leak into the number (round 1's per-launch event
# brackets did: 0.45 on the driver's box vs 0.63 on the builder's).  A torch.profiler (CUPTI) pass over graph-replayed steps
# gives the same kernels' durations INSIDE the step as a cross-check.
# --------------------------------------------------------------------------------------------------------------------
def _fwd_counts(M, K, N, r, save_h):
    flops = 2.0 * M * K * N + 2.0 * M * r * (K + N)
    byts = 2.0 * M * (K + N) + 2.0 * (K * N + r * (K + N)) + (2.0 * M * r if (save_h and r) else 0.0)
    return flops, byts


class LaunchTape:
    """Records ops.lora_linear_fwd / lora_linear_fwd_grouped / lora_linear_bwd calls (inputs are kept alive by reference, outputs
    are re-allocated inside the graph's private pool at capture time)."""

    KINDS = ("gemm_plain", "gemm_fwd", "bwd_dx_wgrad")

    def __init__(self, ops):
        self.ops = ops
        self.records = []   # (kind, fn, args, kwargs, flops, bytes, shape, n_gemm_launches)
        self._fwd, self._bwd, self._grp = ops.lora_linear_fwd, ops.lora_linear_bwd, ops.lora_linear_fwd_grouped
        self._bwd_dx, self._wg_batch = ops.lora_linear_bwd_dx, ops.lora_wgrad_batch

    def __enter__(self):
        ops = self.ops

        def fwd(x, w, bias, down, up, scale, tokens, save_h=False, out=None, residual=None):
            M, K = x.shape
            N = w.shape[0]
            r = 0 if down is None else down.shape[0]
            fl, by = _fwd_counts(M, K, N, r, save_h)
            if residual is not None:
                by += 2.0 * M * N                  # the residual stream read by the epilogue (the add it replaces is not counted as FLOPs)
            self.records.append(("gemm_fwd" if r else "gemm_plain", self._fwd, (x, w, bias, down, up, scale, tokens),
                                 {"save_h": save_h, "residual": residual}, fl, by, (M, K, N, r), 1))
            return self._fwd(x, w, bias, down, up, scale, tokens, save_h=save_h, out=out, residual=residual)

        def bwd(gy, x, w_t, down_t, up_t, scale, h, g_down, g_up, g_scale, tokens):
            M, N = gy.shape      # N = dout
            K = x.shape[1]       # K = din
            r = h.shape[1]
            dx = w_t is not None
            fl = (2.0 * M * K * N if dx else 0.0) + 2.0 * M * r * N + (2.0 * M * r * K if dx else 0.0) + 2.0 * M * r * (K + N)
            by = 2.0 * M * (N + K + r) + (2.0 * M * K if dx else 0.0) + 2.0 * (K * N + r * (K + N)) + 8.0 * r * (K + N)
            self.records.append(("bwd_dx_wgrad", self._bwd, (gy, x, w_t, down_t, up_t, scale, h, g_down, g_up, g_scale, tokens), {},
                                 fl, by, (M, K, N, r), 1))
            return self._bwd(gy, x, w_t, down_t, up_t, scale, h, g_down, g_up, g_scale, tokens)

        def grouped(x, projections, scale, tokens, save_h=False):
            # one launch carrying several projections of the same rows: the algorithmic counts are the per-projection sums
            # (x is counted once per projection, as the reference's separate module calls read it)
            M, K = x.shape
            r = 0 if projections[0][2] is None else projections[0][2].shape[0]
            fl = by = 0.0
            for w, _b, _d, _u in projections:
                f1, b1 = _fwd_counts(M, K, w.shape[0], r, save_h)
                fl += f1; by += b1
            self.records.append(("gemm_fwd" if r else "gemm_plain", self._grp, (x, list(projections), scale, tokens), {"save_h": save_h},
                                 fl, by, (M, K, sum(w.shape[0] for w, *_ in projections), r), 1))
            return self._grp(x, projections, scale, tokens, save_h=save_h)

        def bwd_dx(gy, w_t, down_t, up_t, scale, h, g_scale, tokens):
            # first half of a layer's backward: dX = G W + ((G Up) (.) s) Dn with dH / Hs / dscale side outputs
            M, N = gy.shape
            K, r = down_t.shape
            dx = w_t is not None
            fl = (2.0 * M * K * N if dx else 0.0) + 2.0 * M * r * N + (2.0 * M * r * K if dx else 0.0)
            by = 2.0 * M * (N + r) + (2.0 * M * K if dx else 0.0) + 2.0 * (K * N + r * (K + N)) + 4.0 * M * r
            self.records.append(("bwd_dx_wgrad", self._bwd_dx, (gy, w_t, down_t, up_t, scale, h, g_scale, tokens), {}, fl, by, (M, K, N, r), 1))
            return self._bwd_dx(gy, w_t, down_t, up_t, scale, h, g_scale, tokens)

        def wg_batch(jobs):
            # second half, several layers per launch: dUp += G^T Hs, dDn += dH^T X
            fl = by = 0.0
            for gy, x, _ws, g_down, g_up in jobs:
                M, N = gy.shape
                r, K = g_down.shape
                fl += 2.0 * M * r * (K + N)
                by += 2.0 * M * (N + K + 2 * r) + 8.0 * r * (K + N)
            self.records.append(("bwd_dx_wgrad", self._wg_batch, (list(jobs),), {}, fl, by, (0, 0, 0, len(jobs)), 0))
            return self._wg_batch(jobs)

        ops.lora_linear_fwd, ops.lora_linear_bwd, ops.lora_linear_fwd_grouped = fwd, bwd, grouped
        ops.lora_linear_bwd_dx, ops.lora_wgrad_batch = bwd_dx, wg_batch
        return self

    def __exit__(self, *exc):
        self.ops.lora_linear_fwd, self.ops.lora_linear_bwd, self.ops.lora_linear_fwd_grouped = self._fwd, self._bwd, self._grp
        self.ops.lora_linear_bwd_dx, self.ops.lora_wgrad_batch = self._bwd_dx, self._wg_batch

    def _graph(self, recs):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for rec in recs:                       # warm-up on the capture stream (allocator, attribute opt-ins)
                rec[1](*rec[2], **rec[3])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for rec in recs:
                rec[1](*rec[2], **rec[3])
        return g

    @staticmethod
    def _time(g, min_ms=400.0, max_reps=400):
        """ms per replay: CUDA events on the launching stream around back-to-back replays (>= 0.4 s: sustained clocks)."""
        for _ in range(2):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        reps = int(min(max_reps, max(5, min_ms / max(e0.elapsed_time(e1), 1e-3))))
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, reps

    def summary(self, per_shape: bool = False):
        """{kind: launches, ms per step, TFLOP/s, GB/s} from graph replays that hold only that kind's launches (each replay streams
        every launch's real operands: 12 GB per forward replay, far beyond the 126 MB L2).  per_shape: one graph per (kind, shape)
        over all of its launches of the step -- shapes with a single launch per step then run L2-warm, noted in the table."""
        out, shapes = {}, []
        for kind in self.KINDS:
            recs = [r for r in self.records if r[0] == kind]
            if not recs:
                continue
            g = self._graph(recs)
            ms, reps = self._time(g)
            fl, by = sum(r[4] for r in recs), sum(r[5] for r in recs)
            out[kind] = {"launch_groups_per_step": len(recs), "ms_per_step": ms, "replays_timed": reps, "tflops": fl / ms / 1e9,
                         "gbs": by / ms / 1e6, "flops_per_step": fl, "bytes_per_step": by}
            del g
            if per_shape:
                by_shape = {}
                for r in recs:
                    by_shape.setdefault(r[6], []).append(r)
                for shape, rs in sorted(by_shape.items()):
                    g = self._graph(rs)
                    ms_s, _ = self._time(g, min_ms=60.0)
                    shapes.append({"kind": kind, "M": shape[0], "K": shape[1], "N": shape[2], "r": shape[3], "calls_per_step": len(rs),
                                   "us_per_call": ms_s / len(rs) * 1e3, "tflops": sum(r[4] for r in rs) / ms_s / 1e9,
                                   "l2_warm": sum(r[5] for r in rs) < 2.5e8})
                    del g
        return out, shapes


def cupti_in_step(step_fn, tape_records, steps=2):
    """Durations of this library's GEMM kernels INSIDE graph-replayed steps, from CUPTI activity records (torch.profiler): the
    i-th `lora_gemm_kernel` of a step is the i-th taped launch (same issue order), which labels it plain / fused-forward / dX."""
    try:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                step_fn()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()]
        evs.sort(key=lambda e: e.time_range.start)
        gemm = [e for e in evs if "lora_gemm_kernel" in e.name]
        wgrad = [e for e in evs if "lora_wgrad_kernel" in e.name]
        total_us = sum(e.device_time for e in evs if "Memcpy" not in e.name and "Memset" not in e.name)
        tape_records = [r for r in tape_records if r[7] > 0]          # records that launch the GEMM kernel (not the wgrad batches)
        n = len(tape_records)
        if n == 0 or len(gemm) != n * steps:
            return {"error": f"{len(gemm)} lora_gemm_kernel records for {n} taped launches x {steps} steps"}
        agg = {}
        for i, e in enumerate(gemm):
            agg[tape_records[i % n][0]] = agg.get(tape_records[i % n][0], 0.0) + e.device_time
        out = {k: round(v / steps / 1e3, 3) for k, v in agg.items()}          # ms per step
        out["wgrad_ms"] = round(sum(e.device_time for e in wgrad) / steps / 1e3, 3)
        out["all_kernels_ms"] = round(total_us / steps / 1e3, 3)
        return out
    except Exception as e:  # the cross-check must never cost the measurement
        return {"error": repr(e)[:200]}


def gpu_eager_baseline(model, B, dev, steps=5, warmup=3):
    """The bar SURVEY.md 0.1 names, on the SAME B200: the reference's UNFUSED op sequence (utils/lora_modules.py:9-62 restated in
    oracle/: Linear, down, diag_embed + bmm, up, add -- each its own cuBLAS call) on PyTorch-eager with library GroupNorm / LayerNorm /
    GELU, bf16 autocast over fp32 LoRA master weights as `accelerate --mixed_precision bf16` runs it (train/ppft_train.py:569-581),
    clip_grad_norm_ + torch.optim.AdamW, same batch size.  None of this repository's kernels is on this path."""
    from aqualora_b200 import lora_modules, ppft
    from aqualora_b200 import unet as unet_mod
    from aqualora_b200.unet import UNetConfig, lora_target_keys
    from oracle import lora_oracle as O
    from oracle.patch import patch_with_oracle

    cfg = UNetConfig.sd15(64) if model == "sd15" else UNetConfig.sd21(96)
    unet_mod.LIBRARY_GLUE = True
    try:
        unet = ppft.build_unet(cfg, dev, seed=0)
        layers = lora_modules.inject_lora(unet, lora_target_keys(unet), RANK_R)
        g = torch.Generator().manual_seed(1)
        params = []
        for _, _, l in layers:
            l.up.weight.data.copy_(torch.randn(l.up.weight.shape, generator=g) * 0.02)
            for p in (l.down.weight, l.up.weight):
                p.requires_grad_(True)
                params.append(p)
        patch_with_oracle(unet)
        emb = O.mapper_init(BITS, RANK_R, generator=torch.Generator().manual_seed(5)).to(dev).requires_grad_(True)
        opt = torch.optim.AdamW([{"params": params}, {"params": [emb]}], lr=1e-4, weight_decay=1e-2)
        ac = ppft.scaled_linear_alphas_cumprod().to(dev)
        v_pred = model != "sd15"
        batches = [tuple(x.to(dev) for x in synth_batch(B, cfg, 1234 + i)) for i in range(2)]
        bf = torch.bfloat16

        def step(i):
            lat, noise, t, ctx, msg = batches[i % 2]
            with torch.autocast("cuda", dtype=bf):
                scale = O.mapper_forward(msg, emb).to(bf)
                wm = (torch.randn_like(lat) * 0.004).to(bf)           # stands in for the (no_grad) encoder residual
                noisy = ppft.add_noise(ac, lat.to(bf), noise.to(bf), t)
                noisy_wm = ppft.add_noise(ac, lat.to(bf) + wm, noise.to(bf), t)
                clean = unet(noisy, t, ctx.to(bf), cross_attention_kwargs={"scale": torch.zeros_like(scale)}).sample.detach()
                pred = unet(noisy_wm, t, ctx.to(bf), cross_attention_kwargs={"scale": scale}).sample
                if v_pred:
                    pred = ppft.velocity_to_epsilon(ac, pred, noisy_wm, t)
                    clean = ppft.velocity_to_epsilon(ac, clean, noisy, t)
                loss = torch.nn.functional.mse_loss(pred.float(), clean.float())
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad()
            return loss

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(B / ms * 1e3, 3), "unit": UNIT, "ms_per_step": round(ms, 3), "per_gpu_batch": B, "steps": steps,
                "what": "reference op sequence (unfused LoRA: Linear + down + diag_embed/bmm + up + add) on PyTorch eager, cuBLAS / cuDNN, "
                        "bf16 autocast, library norms, torch AdamW; same B200, same batch, no CUDA graph"}
    finally:
        unet_mod.LIBRARY_GLUE = False


# --------------------------------------------------------------------------------------------------------------------
# the CUDA arm
# --------------------------------------------------------------------------------------------------------------------
_JSON_OUT = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner to fd 1) are re-pointed at stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


# dram__bytes_read.sum + dram__bytes_write.sum of the fused forward kernel, one ncu launch per shape of the SD1.5 B = 16 step
# (profiles/r01_gemm_dram_traffic_v13.txt), weighted by the shape's launches per step: 7 890 MB over 129 launches
GEMM_FWD_DRAM_TRAFFIC_PER_LAUNCH = round(7890e6 / 129)


def run_cuda(args):
    _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this repository has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from aqualora_b200 import _lib, build, ops, ppft
    from aqualora_b200.models import MapperNet, SecretEncoder
    from aqualora_b200.unet import UNetConfig

    if not _lib.LIB_PATH.exists():
        if rank == 0:
            build.build()
        if world > 1:
            dist.barrier()
    lib = _lib.load()

    B = args.batch
    cfg = UNetConfig.sd15(64) if args.model == "sd15" else UNetConfig.sd21(96)
    unet = ppft.build_unet(cfg, dev, seed=0)
    torch.manual_seed(5)
    mapper = MapperNet(BITS, RANK_R)
    pcfg = ppft.PPFTConfig(rank=RANK_R, msg_bits=BITS, max_train_steps=10_000,
                           prediction_type="epsilon" if args.model == "sd15" else "v_prediction")
    trainer = ppft.PPFTTrainer(unet, pcfg, mapper.bit_embeddings.weight.data, dev, lora_up_std=0.02, seed=1)
    torch.manual_seed(0)
    enc = SecretEncoder(BITS).to(dev)
    torch.nn.init.normal_(enc.secret_scaler[5].weight, std=0.02)   # the reference zero-inits this conv; trained value assumed

    def to_dev(batch, non_blocking=False):
        lat, noise, t, ctx, msg = batch
        f = lambda x: x.to(dev, non_blocking=non_blocking)
        return f(lat), f(noise), f(t), f(ctx), f(msg)

    def fwd_bwd_from_device(lat, noise, t, ctx, msg):
        # train/ppft_train.py:994-996: secret residual from the encoder (no_grad), scaled like the latents
        wm = enc(lat, msg)[1] * pcfg.scaling_factor      # sec_encoder(latents, msg)[1]: the residual, resized to the latent grid
        bf = torch.bfloat16
        return trainer.forward_backward(lat.to(bf), wm.to(bf), noise.to(bf), t, ctx.to(bf), msg)

    def eager_step(lat, noise, t, ctx, msg):
        loss = fwd_bwd_from_device(lat, noise, t, ctx, msg)
        trainer.optimizer_step()
        return loss

    graph_state = {"on": False, "launches": 0, "why": "disabled by --no-graph"}

    def step_from_device(lat, noise, t, ctx, msg):
        if graph_state["on"]:
            return trainer.step_graphed(lat, noise, t, ctx, msg)
        return eager_step(lat, noise, t, ctx, msg)

    n_pool = 4
    host = [tuple(x.pin_memory() for x in synth_batch(B, cfg, 1234 + i + 1000 * rank)) for i in range(n_pool)]
    resident = [to_dev(b) for b in host]
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), then the forward + backward recorded as one CUDA graph; more warm-up through the graph -----
    for i in range(3):
        loss = eager_step(*resident[i % n_pool])
    barrier()
    if not args.no_graph:
        try:
            graph_state["launches"] = trainer.capture(fwd_bwd_from_device, resident[0])
            graph_state["on"], graph_state["why"] = True, ""
        except Exception as e:   # a capture failure must not cost the measurement: fall back to eager issue and say so
            graph_state["why"] = repr(e)[:200]
            print(f"[bench] CUDA-graph capture failed, running eagerly: {e!r}", file=sys.stderr, flush=True)
            torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        loss = step_from_device(*resident[i % n_pool])
    barrier()

    # ---- timed region 1: inputs resident in HBM -------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.aq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        loss = step_from_device(*resident[i % n_pool])
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = lib.aq_launch_count() - launches0 + (args.steps * graph_state["launches"] if graph_state["on"] else 0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    final_loss = float(loss)

    # ---- timed region 2: end to end from pinned host memory -------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        if graph_state["on"]:
            for dst, src in zip(trainer.static_inputs, host[i % n_pool]):     # pinned host -> the graph's input buffers
                dst.copy_(src, non_blocking=True)
            loss_host = float(trainer.step_graphed(copy_inputs=False))       # D2H read of the step's result
        else:
            batch = to_dev(host[i % n_pool], non_blocking=True)
            loss_host = float(step_from_device(*batch))
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = ms2.item()

    # ---- roofline pass (after the timed regions; rank 0 only, no collective inside) ---------------------------------
    roof = None
    extra = {}
    if rank == 0:
        pk = peaks()
        with LaunchTape(ops) as tape:
            fwd_bwd_from_device(*resident[0])          # ONE eagerly issued forward + backward, recorded (no optimizer step)
        torch.cuda.synchronize()
        kinds, shapes = tape.summary(per_shape=bool(args.shapes_out))
        in_step = None
        if graph_state["on"]:
            in_step = cupti_in_step(lambda: trainer._graph.replay(), tape.records)
        trainer.state.grad.zero_()                     # the probes accumulated weight gradients that belong to no step
        if trainer.g_scale is not None:
            trainer.g_scale.zero_()
        dom = kinds.get("gemm_fwd")
        if dom:
            peak = pk["bf16_tflops_sustained"]         # back-to-back replays for >= 0.4 s: the power-capped regime
            n_l = dom["launch_groups_per_step"]
            roof = {"bound": "tensor", "kernel": "aq::lora_gemm_kernel (fused base GEMM + watermark LoRA, forward launches)",
                    "achieved": round(dom["tflops"], 1), "peak": peak, "unit": "TFLOP/s", "frac": round(dom["tflops"] / peak, 4),
                    "peak_kind": f"bf16_tflops_sustained of {pk['source']}",
                    "traffic": (GEMM_FWD_DRAM_TRAFFIC_PER_LAUNCH if args.model == "sd15" and B == PER_GPU_BATCH else None),
                    "traffic_note": "DRAM bytes (read + write) per forward launch, averaged over the launches of a step: ncu "
                                    "dram__bytes_* of every forward shape, cold caches (profiles/r01_gemm_dram_traffic_v13.txt); "
                                    "algorithmic bytes per launch = bytes_per_step / launches (outputs largely stay in the 126 MB L2)",
                    "launches_per_step": n_l, "kernel_ms_per_step": round(dom["ms_per_step"], 3),
                    "avg_launch_us": round(dom["ms_per_step"] / n_l * 1e3, 2), "flops_per_launch": dom["flops_per_step"] / n_l,
                    "flops_per_step": dom["flops_per_step"], "frac_of_burst": round(dom["tflops"] / pk["bf16_tflops"], 4),
                    "in_step_cupti_ms": None if not in_step else in_step.get("gemm_fwd"),
                    "frac_in_step_cupti": (round(dom["flops_per_step"] / in_step["gemm_fwd"] / 1e9 / peak, 4)
                                           if in_step and in_step.get("gemm_fwd") else None),
                    "method": "device-only: the forward launches of one recorded step (real shapes, operands, pointers) re-issued back to "
                              f"back in a CUDA graph; CUDA events on the launching stream around {dom['replays_timed']} replays (kernel-to-kernel "
                              "gaps included, no host in the timed region; every replay streams 12 GB of operands, L2 126 MB); "
                              "in_step_cupti_ms = the same kernels' summed CUPTI durations inside graph-replayed training steps"}
        extra = {"kernel_groups": {k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in kinds.items()},
                 "in_step_cupti": in_step}
        if args.shapes_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.shapes_out)), exist_ok=True)
            json.dump({"kinds": kinds, "shapes": shapes, "in_step_cupti": in_step}, open(args.shapes_out, "w"), indent=1)
        del tape

    # ---- the same-box GPU baseline: reference op sequence on PyTorch eager (rank 0, N = 1 only) --------------------
    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        try:
            eager = gpu_eager_baseline(args.model, B, dev)
        except Exception as e:  # must not cost the headline
            eager = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_arm(args.model, steps=4, warmup=1, budget_s=45.0)     # ~11-14 s of CPU work beside the GPU number

    if rank == 0:
        gb = B * world
        ms_step = ms_total / args.steps
        line = {
            "metric": METRIC, "value": round(gb / ms_step * 1e3, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.model, cfg.sample_size),
                       "per_gpu_batch": B, "global_batch": gb, "parallelism": f"dp{world}",
                       "l2": "each step streams > 126 MB (1.7 GB bf16 U-Net weights + activations), inputs rotate over 4 batches",
                       "cuda_graph": ("forward + backward replayed as one CUDA graph; all-reduce, clip and AdamW issued eagerly"
                                      if graph_state["on"] else f"off ({graph_state['why']})"),
                       "final_loss": final_loss},
            "e2e": {"value": round(gb / (e2e_ms / args.steps) * 1e3, 3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_ms / args.steps, 3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": eager,
        }
        line.update(extra)
        if world == 1 and not args.no_secondary:
            # BASELINE.json configs[3] (decode throughput + bit agreement) rides along as a secondary record; the trainer is
            # released first so the decoder's workspace and image pool fit beside nothing else
            del trainer, unet, resident, host
            torch.cuda.empty_cache()
            try:
                d = measure_decode(args.decode_images)
                line["secondary"] = {k: d[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "config", "gpu_launches",
                                                       "bit_agreement", "roofline", "cpu_baseline")}
            except Exception as e:  # the headline line must survive a failure of the secondary workload
                line["secondary"] = {"metric": "decode_images_per_sec", "error": repr(e)[:300]}
        print(json.dumps(line), file=_claim_stdout(), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------------
# the reference arm: the reference's op sequence (oracle restatement, PyTorch eager) on the host CPU
# --------------------------------------------------------------------------------------------------------------------
def cpu_reference_arm(model: str, steps: int, warmup: int, budget_s: float | None = None):
    """PPFT step exactly as train/ppft_train.py:987-1068 issues it (clean forward WITH the zero-scale LoRA branch, watermarked
    forward, MSE, backward, clip, AdamW), with the reference's unfused LoRA forwards (oracle/lora_oracle.py, pinned to the
    reference's own utils/lora_modules.py by tests/golden) on the PyTorch-eager U-Net, fp32, B = 1, all host threads."""
    from aqualora_b200 import lora_modules, ppft
    from aqualora_b200.unet import UNetConfig, lora_target_keys
    from oracle import lora_oracle as O
    from oracle import models_oracle as MO
    from oracle.patch import patch_with_oracle

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = UNetConfig.sd15(64) if model == "sd15" else UNetConfig.sd21(96)
    B = 1
    unet = ppft.build_unet(cfg, "cpu", dtype=torch.float32, seed=0)
    layers = lora_modules.inject_lora(unet, lora_target_keys(unet), RANK_R)
    g = torch.Generator().manual_seed(1)
    params = []
    for _, _, l in layers:
        l.up.weight.data.copy_(torch.randn(l.up.weight.shape, generator=g) * 0.02)
        for p in (l.down.weight, l.up.weight):
            p.requires_grad_(True)
            params.append(p)
    patch_with_oracle(unet)
    emb = O.mapper_init(BITS, RANK_R, generator=torch.Generator().manual_seed(5)).requires_grad_(True)
    torch.manual_seed(0)
    enc_sd = {"secret_scaler.0.weight": torch.randn(1024, BITS) * BITS ** -0.5, "secret_scaler.0.bias": torch.zeros(1024),
              "secret_scaler.5.weight": torch.randn(4, 4, 3, 3) * 0.02, "secret_scaler.5.bias": torch.zeros(4)}
    opt = torch.optim.AdamW([{"params": params}, {"params": [emb]}], lr=1e-4, weight_decay=1e-2)
    ac = ppft.scaled_linear_alphas_cumprod()
    v_pred = model != "sd15"

    def step(i):
        lat, noise, t, ctx, msg = synth_batch(B, cfg, 1234 + i)
        scale = O.mapper_forward(msg, emb)
        with torch.no_grad():
            wm = MO.secret_encoder_forward(lat, msg, enc_sd)[1] * 0.18215
        noisy = ppft.add_noise(ac, lat, noise, t)
        noisy_wm = ppft.add_noise(ac, lat + wm, noise, t)
        clean = unet(noisy, t, ctx, cross_attention_kwargs={"scale": torch.zeros_like(scale)}).sample.detach()
        pred = unet(noisy_wm, t, ctx, cross_attention_kwargs={"scale": scale}).sample
        if v_pred:
            pred = ppft.velocity_to_epsilon(ac, pred, noisy_wm, t)
            clean = ppft.velocity_to_epsilon(ac, clean, noisy, t)
        loss = torch.nn.functional.mse_loss(pred.float(), clean.float())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad()
        return float(loss.detach())

    for i in range(warmup):
        step(i)
    t0 = time.time()
    done = 0
    for i in range(steps):
        step(warmup + i)
        done += 1
        if budget_s is not None and time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    return {"value": round(B * done / dt, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} PPFT step(s) at B=1 (full {model} U-Net, fp32, PyTorch eager + the reference's unfused LoRA forwards), "
                      f"{dt:.1f} s on {cores} threads", "seconds": round(dt, 2), "steps_done": done}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    res = cpu_reference_arm(args.model, steps=args.steps, warmup=min(args.warmup, 1), budget_s=args.budget)
    ms_step = res["seconds"] / max(res["steps_done"], 1) * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": res["steps_done"],
            "warmup": args.warmup, "ms_per_step": round(ms_step, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.model, 64 if args.model == "sd15" else 96),
                       "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus, "parallelism": f"dp{args.gpus}",
                       "reference_sample": "each timed step is a bounded sample of that workload: ONE image (B = 1) through the full-size "
                                           "step on the host cores, fp32 (the reference's CPU precision); images/s = images done / time",
                       "note": "reference op sequence = the reference's unfused LoRA forwards (oracle/lora_oracle.py, bit-identical to "
                               "utils/lora_modules.py on tests/golden) on the U-Net harness that tests/test_unet_golden.py pins to the "
                               "reference's vendored scripts/lib/original_unet.py; /root/reference itself does not exist on the GPU box"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# secondary workload (BASELINE.json configs[3]): decode throughput, noise_layers + Decoder, bit agreement vs the oracle
# --------------------------------------------------------------------------------------------------------------------
DECODE_BYTES_PER_IMAGE_FP32 = 404.4e6     # 101.09 M activation elements with ideal per-layer fusion, fp32
NOISE_P = [0.4, 0.1, 0.2, 0.05, 0.1, 0.15]


# dram__bytes_read.sum + dram__bytes_write.sum over the 93 decoder launches of one 64-image batch (ncu launch list
# profiles/r02_decoder_launches_v23.txt: 13.77 GB read + 9.73 GB written), per image
DECODE_DRAM_TRAFFIC_PER_IMAGE = (13.767e9 + 9.729e9) / 64


def run_decode(args):
    """`--workload decode` (BASELINE.json configs[3]).  Under torchrun the images are sharded over the ranks (independent units, no
    data-path collective: SURVEY.md 8(e)); the only exchange is the final reduction of the timing (max) and of the bit counts."""
    out = _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        print(json.dumps(measure_decode(args.images)), file=out, flush=True)
        return
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dist.barrier()
    line = measure_decode(max(64, args.images // world), device_index=local_rank, shard=rank)
    dev = torch.device("cuda", local_rank)
    t_max = torch.tensor([line["ms_per_step"] * line["steps"]], device=dev, dtype=torch.float64)
    counts = torch.tensor([line["steps"] * 64, line["bit_agreement"]["agree"], line["bit_agreement"]["total"],
                           line["bit_agreement"]["below_fp32_noise_floor"], line["gpu_launches"]], device=dev, dtype=torch.int64)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(counts)
    if rank == 0:
        imgs, agree, total, undec, launches = (int(v) for v in counts.tolist())
        line.update({"value": round(imgs / t_max.item() * 1e3, 1), "n_gpus": world, "gpu_launches": launches})
        line["bit_agreement"].update({"agree": agree, "total": total, "below_fp32_noise_floor": undec})
        line["config"]["parallelism"] = f"images sharded over {world} ranks, no data-path collective"
        print(json.dumps(line), file=out, flush=True)
    dist.destroy_process_group()


def measure_decode(images, cpu_check=True, device_index=0, shard=0, check_all=True):
    """`--workload decode`: N synthetic 512x512 images in batches of 64; per batch one noise layer drawn with
    p = [.4, .1, .2, .05, .1, .15] (train/latent_wm_pretrain.py:188) from numpy default_rng(7), then the EfficientNet-B1
    decoder; bits checked against the CPU oracle on a bounded sample."""
    import numpy as np

    from aqualora_b200 import _lib, noise_layers
    from aqualora_b200.decoder import SecretDecoder
    from oracle import models_oracle as MO
    from oracle import noise_oracle as NO

    dev = torch.device("cuda", device_index)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    sd, _ = MO.make_decoder_state(BITS, seed=0)
    dec = SecretDecoder(BITS)
    dec.load_state_dict(sd)
    dec = dec.to(dev).eval()
    bs = 64
    n_batches = max(1, images // bs)
    names = ["Jpeg", "CropandResize", "GaussianBlur", "GaussianNoise", "ColorJitter"]
    noiser = noise_layers.Noiser(names, NOISE_P, dev, rng=np.random.default_rng(7))
    pool = [noise_layers.unit_noise((bs, 3, 512, 512), seed=7, offset=(4 * shard + i) * bs * 3 * 512 * 512 // 4, device=dev).clamp_(-3, 3) / 3
            for i in range(4)]                                   # 4 x 201 MB of images: larger than L2 (a different slice per shard)

    drawn = []         # (layer index, parameters) the Noiser drew for every timed batch: the checker replays them

    def batch_step(i, record=False):
        img = noiser([pool[i % 4], None])[0]
        bits_i = dec.decode_bits(img)
        if record:
            layer = noiser.noise_layers[noiser.last_layer]
            drawn.append((noiser.last_layer, dict(getattr(layer, "last_params", {})), bits_i))
        return bits_i

    for i in range(3):
        batch_step(i)
    torch.cuda.synchronize()
    n0 = lib.aq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_batches):
        bits = batch_step(i, record=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.aq_launch_count() - n0
    # decoder alone (the HBM roofline the north star quotes)
    e0.record()
    for i in range(10):
        dec.decode_bits(pool[i % 4])
    e1.record()
    torch.cuda.synchronize()
    dec_ms = e0.elapsed_time(e1) / 10

    # ---- parity, every image (BASELINE configs[3]: "bit-acc vs ref" over all images) ---------------------------------------
    # checker = the oracle restatement itself (oracle/noise_oracle.py + oracle/models_oracle.py: plain torch fp32 ops) executed on the
    # GPU with TF32 off, because the CPU oracle needs ~25 ms per image; it is pinned to the CPU oracle on a 48-image sample below.
    tf32_state = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd_dev = {k: v.to(dev) for k, v in sd.items()}

    def oracle_on(x, li, lp, device_sd):
        if li == 4:
            noise = noise_layers.unit_noise(tuple(x.shape), lp["seed"], lp["offset"], device=dev).to(x.device)
            img = NO.gaussian_noise(x, lp["std"], noise)
        else:
            img = NO.apply_layer(x, li, {k: v for k, v in lp.items()})
        with torch.no_grad():
            return MO.secret_decoder_forward(img, device_sd, BITS)

    agree = total = undecidable = 0
    worst_margin_flip = 0.0
    per_layer = {}
    if check_all:
        for i, (li, lp, bits_i) in enumerate(drawn):
            want = oracle_on(pool[i % 4], li, lp, sd_dev)
            margin = (want[..., 0] - want[..., 1]).abs()
            ok = bits_i.long() == want.argmax(-1)
            dec_ok = margin > 2e-4 * want.abs().max()
            agree += int(ok.sum()); total += ok.numel(); undecidable += int((~dec_ok).sum())
            if not bool(ok.all()):
                worst_margin_flip = max(worst_margin_flip, float((margin[~ok] / want.abs().max()).max()))
            a = per_layer.setdefault(NO.LAYER_NAMES[li], [0, 0])
            a[0] += int(ok.sum()); a[1] += ok.numel()
    # the checker against the CPU oracle: 8 images per noise layer, identical parameters
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    t_cpu = 0.0
    x_cpu = pool[0][:8].cpu()
    pin_diff, pin_bits_equal, s_agree, s_total = 0.0, True, 0, 0
    for li, layer in enumerate(noiser.noise_layers):
        xin = pool[0][:8].clone()
        out = layer([xin, None])[0]
        lp = dict(getattr(layer, "last_params", {}))
        t0 = time.time()
        want = oracle_on(x_cpu, li, lp, sd)
        t_cpu += time.time() - t0
        want_gpu = oracle_on(pool[0][:8], li, lp, sd_dev).cpu()
        pin_diff = max(pin_diff, float((want_gpu - want).abs().max() / want.abs().max()))
        margin = (want[..., 0] - want[..., 1]).abs()
        dec_ok = margin > 2e-4 * want.abs().max()
        pin_bits_equal = pin_bits_equal and bool((want_gpu.argmax(-1) == want.argmax(-1))[dec_ok].all())
        got_bits = dec.decode_bits(out).cpu().long()
        s_agree += int((got_bits == want.argmax(-1)).sum()); s_total += want.argmax(-1).numel()
        if not check_all:
            agree += int((got_bits == want.argmax(-1)).sum()); total += want.argmax(-1).numel(); undecidable += int((~dec_ok).sum())
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_state
    pk = peaks()
    imgs = n_batches * bs
    dec_rate = bs / dec_ms * 1e3
    line = {"metric": "decode_images_per_sec", "value": round(imgs / ms * 1e3, 1), "unit": "images/s", "n_gpus": 1, "steps": n_batches,
            "warmup": 3, "ms_per_step": round(ms / n_batches, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{imgs} synthetic 512x512 images, batches of {bs}: one noise layer per batch (p={NOISE_P}) + "
                                   "EfficientNet-B1 decoder (random-init, BN stats randomised), bits = argmax", "l2": "4 x 201 MB input pool"},
            "gpu_launches": int(launches),
            "bit_agreement": {"agree": agree, "total": total, "below_fp32_noise_floor": undecidable,
                              "largest_relative_margin_of_a_disagreeing_bit": worst_margin_flip, "per_layer": per_layer,
                              "sample": (f"EVERY image of the timed run ({imgs} images x {BITS} bits), checker = the oracle restatement run on the GPU "
                                         "in fp32 (TF32 off) with the layer parameters the timed run drew" if check_all else
                                         "8 images x 6 noise layers through the CPU oracle with identical layer parameters"),
                              "checker_pinned_to_cpu_oracle": {"images": 48, "max_rel_logit_diff": pin_diff, "bits_equal_where_decidable": pin_bits_equal,
                                                               "cuda_path_vs_cpu_oracle_bits": [s_agree, s_total]}},
            "roofline": {"bound": "hbm", "kernel": "decoder chain (csrc/decoder.cu), decoder-only loop", "achieved": round(DECODE_BYTES_PER_IMAGE_FP32 * dec_rate / 1e9, 1),
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(DECODE_BYTES_PER_IMAGE_FP32 * dec_rate / 1e9 / pk["hbm_gbs"], 4),
                         "traffic": round(DECODE_DRAM_TRAFFIC_PER_IMAGE), "traffic_note": "DRAM bytes per image, ncu launch list of one 64-image batch (profiles/r02_decoder_launches_v23.txt)",
                         "decoder_images_per_sec": round(dec_rate, 1),
                         "algorithmic_bytes_per_image": DECODE_BYTES_PER_IMAGE_FP32},
            "cpu_baseline": {"value": round(48 / t_cpu, 3), "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"48 images (8 per noise layer) through oracle noise layer + torch EfficientNet-B1 restatement, {t_cpu:.1f} s"}}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--model", default="sd15", choices=["sd15", "sd21"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager (unfused reference op sequence) GPU baseline")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel eagerly instead of replaying the captured forward + backward")
    ap.add_argument("--budget", type=float, default=240.0, help="--impl reference: stop after this many seconds of timed CPU steps")
    ap.add_argument("--shapes-out", default=None, help="write the per-shape kernel table (JSON) here")
    ap.add_argument("--workload", default="ppft", choices=["ppft", "decode"], help="ppft = the headline metric; decode = configs[3]")
    ap.add_argument("--images", type=int, default=10_000, help="--workload decode: number of images")
    ap.add_argument("--decode-images", type=int, default=2048, help="images of the secondary decode record on the ppft line (N = 1)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary decode record")
    args = ap.parse_args()
    if args.workload == "decode":
        run_decode(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
