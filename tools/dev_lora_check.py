"""Developer harness for the LoRA kernels on a real B200: correctness against a torch fp32 reference plus CUDA-event
timings.  Each case runs in its own subprocess (a device trap poisons the CUDA context) under a timeout.

    python tools/dev_lora_check.py            # run every case
    python tools/dev_lora_check.py --case fwd:M=1024,K=320,N=320,r=64,bn=160
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    "plain:M=128,K=64,N=64,bn=64",
    "plain:M=256,K=128,N=128,bn=128",
    "plain:M=1024,K=320,N=320,bn=160",
    "plain:M=1000,K=328,N=200,bn=0",
    "plain:M=128,K=320,N=320,bn=160",
    "plain:M=384,K=640,N=640,bn=192",
    "plain:M=40000,K=320,N=1280,bn=0",
    "fwd:M=128,K=64,N=64,r=64,tok=128,bn=64",
    "fwd:M=1024,K=320,N=320,r=64,tok=256,bn=160",
    "fwd:M=1232,K=768,N=320,r=64,tok=77,bn=160",
    "fwd:M=2048,K=640,N=5120,r=16,tok=1024,bn=0",
    "fwd:M=4096,K=1280,N=1280,r=64,tok=256,bn=192",
    "fwd:M=4096,K=1280,N=1280,r=64,tok=256,bn=160,grp=1",
    "fwd:M=4096,K=1280,N=1280,r=64,tok=256,bn=160,grp=3",
    "fwd:M=4096,K=1280,N=1280,r=64,tok=256,bn=128,grp=10",
    "fwd:M=384,K=320,N=320,r=8,tok=128,bn=0",
    "fwd:M=65536,K=320,N=320,r=64,tok=4096,bn=0",
    "fwd:M=20000,K=640,N=5120,r=64,tok=1000,bn=0",
    "wgrad:M=1024,I=320,J=64,t=0",
    "wgrad:M=5000,I=1280,J=64,t=1",
    "wgrad:M=1232,I=768,J=16,t=1",
    "bwd:M=1024,K=320,N=320,r=64,tok=256",
    "bwd:M=1232,K=768,N=320,r=64,tok=77,nodx=1",
    "bwd:M=2048,K=640,N=2560,r=32,tok=1024",
    "bwd:M=16384,K=640,N=640,r=64,tok=1024",
    "bwd:M=900,K=320,N=1280,r=64,tok=300",
    "time:M=65536,K=320,N=320,r=64,tok=4096",
    "time:M=65536,K=320,N=2560,r=64,tok=4096",
    "time:M=65536,K=1280,N=320,r=64,tok=4096",
    "time:M=16384,K=640,N=640,r=64,tok=1024",
    "time:M=16384,K=640,N=5120,r=64,tok=1024",
    "time:M=4096,K=1280,N=1280,r=64,tok=256",
    "time:M=4096,K=1280,N=10240,r=64,tok=256",
    "time:M=4096,K=5120,N=1280,r=64,tok=256",
]


def parse(case: str):
    kind, _, rest = case.partition(":")
    kv = {}
    for item in rest.split(","):
        if item:
            k, v = item.split("=")
            kv[k] = int(v)
    return kind, kv


def err_report(name, got, ref, tol):
    import torch

    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel = diff.max().item() / denom
    bad = (diff > tol * denom).nonzero()
    out = {"name": name, "max_abs": diff.max().item(), "ref_max": denom, "rel": rel, "n_bad": int(bad.shape[0]),
           "finite": bool(torch.isfinite(got).all().item())}
    if bad.shape[0]:
        out["first_bad"] = bad[:6].tolist()
        rows = torch.unique(bad[:, 0])
        out["bad_rows_head"] = rows[:16].tolist()
        if bad.shape[1] > 1:
            out["bad_cols_head"] = torch.unique(bad[:, 1])[:16].tolist()
    return out


def run_case(case: str):
    import torch

    from aqualora_b200 import ops

    kind, kv = parse(case)
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)

    def rnd(*shape, s=1.0):
        return (torch.randn(*shape, generator=g) * s).to(dev)

    res = {"case": case}
    if kind in ("plain", "fwd", "time", "bwd"):
        M, K, N = kv["M"], kv["K"], kv["N"]
        r = kv.get("r", 0)
        tok = kv.get("tok", M)
        x = rnd(M, K).bfloat16()
        w = rnd(N, K, s=K ** -0.5).bfloat16()
        b = rnd(N).bfloat16()
        nsamp = (M + tok - 1) // tok
        if r:
            dn = rnd(r, K, s=K ** -0.5).bfloat16()
            up = rnd(N, r, s=0.1).bfloat16()
            sc = (1 + 0.7 * torch.randn(nsamp, r, generator=g)).bfloat16().float().to(dev)
        else:
            dn = up = sc = None
        ops.set_tuning(kv.get("bn", 0), kv.get("grp", 0))

    def ref_fwd():
        y = x.float() @ w.float().t() + b.float()
        h = None
        if r:
            h = (x.float() @ dn.float().t()).bfloat16().float()
            srow = sc.repeat_interleave(tok, dim=0)[:M]
            hs = (h * srow).bfloat16().float()
            y = y + hs @ up.float().t()
        return y, h

    if kind in ("plain", "fwd"):
        y, h = ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=bool(r))
        torch.cuda.synchronize()
        yr, hr = ref_fwd()
        res["y"] = err_report("y", y, yr, 2e-2)
        if r:
            res["h"] = err_report("h", h, hr, 2e-2)
        res["ok"] = res["y"]["n_bad"] == 0 and res["y"]["finite"] and (not r or res["h"]["n_bad"] == 0)
    elif kind == "wgrad":
        M, I, J, t = kv["M"], kv["I"], kv["J"], kv["t"]
        p = rnd(M, I).bfloat16()
        q = rnd(M, J).bfloat16()
        c = torch.zeros((J, I) if t else (I, J), device=dev)
        ops.wgrad_tn(p, q, c, transpose_out=bool(t))
        torch.cuda.synchronize()
        cr = p.float().t() @ q.float()
        if t:
            cr = cr.t()
        res["c"] = err_report("c", c, cr, 2e-3)
        res["ok"] = res["c"]["n_bad"] == 0 and res["c"]["finite"]
    elif kind == "bwd":
        nodx = kv.get("nodx", 0)
        _, h = ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True)
        gy = rnd(M, N, s=0.05).bfloat16()
        w_t = None if nodx else w.t().contiguous()
        g_dn = torch.zeros(r, K, device=dev)
        g_up = torch.zeros(N, r, device=dev)
        g_sc = torch.zeros(nsamp, r, device=dev)
        gx = ops.lora_linear_bwd(gy, x, w_t, dn.t().contiguous(), up.t().contiguous(), sc, h, g_dn, g_up, g_sc, tok)
        torch.cuda.synchronize()
        # fp32 reference with the same bf16 rounding points
        srow = sc.repeat_interleave(tok, dim=0)[:M]
        hf = h.float()
        dhs = (gy.float() @ up.float()).bfloat16().float()
        dh = (dhs * srow).bfloat16().float()
        hs = (hf * srow).bfloat16().float()
        gx_ref = gy.float() @ w.float() + dh @ dn.float()
        gup_ref = gy.float().t() @ hs
        gdn_ref = dh.t() @ x.float()
        prod = dhs * hf
        pad = nsamp * tok - M
        if pad:
            prod = torch.cat([prod, torch.zeros(pad, r, device=dev)])
        gsc_ref = prod.view(nsamp, tok, r).sum(1)
        if not nodx:
            res["gx"] = err_report("gx", gx, gx_ref, 2e-2)
        res["g_up"] = err_report("g_up", g_up, gup_ref, 5e-3)
        res["g_dn"] = err_report("g_dn", g_dn, gdn_ref, 5e-3)
        res["g_sc"] = err_report("g_sc", g_sc, gsc_ref, 5e-3)
        res["ok"] = all(res[k]["n_bad"] == 0 and res[k]["finite"] for k in ("g_up", "g_dn", "g_sc")) and (
            nodx or res["gx"]["n_bad"] == 0)
    elif kind == "time":
        peaks = {"bf16_tflops": 1691.3, "hbm_gbs": 6550.1}
        y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

        def timed(fn, iters=10):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(iters):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            return ts[len(ts) // 2] * 1e-3

        flops_fused = 2.0 * M * K * N + 2.0 * M * r * (K + N)
        t_fused = timed(lambda: ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True, out=y))
        t_plain = timed(lambda: ops.lora_linear_fwd(x, w, b, None, None, None, tok, out=y))
        t_cublas = timed(lambda: torch.addmm(b, x, w.t(), out=y))

        def eager():
            base = torch.nn.functional.linear(x, w, b)
            hh = torch.nn.functional.linear(x, dn)
            hh = hh.view(nsamp, tok, r) @ torch.diag_embed(sc.bfloat16())
            return base + torch.nn.functional.linear(hh.view(M, r), up)

        t_eager = timed(eager) if M % tok == 0 else float("nan")
        res.update({
            "fused_us": t_fused * 1e6, "plain_us": t_plain * 1e6, "cublas_base_us": t_cublas * 1e6, "eager_ref_us": t_eager * 1e6,
            "fused_tflops": flops_fused / t_fused / 1e12, "fused_frac_burst": flops_fused / t_fused / 1e12 / peaks["bf16_tflops"],
            "plain_tflops": 2.0 * M * K * N / t_plain / 1e12, "cublas_tflops": 2.0 * M * K * N / t_cublas / 1e12,
            "bytes_GBs": (2.0 * M * (K + N + r) + 2.0 * (K * N + r * (K + N))) / t_fused / 1e9,
        })
        # sweep group size for the fused kernel
        sweep = {}
        for bn in (160, 192, 128):
            for grp in (1, 2, 4, 8, 16):
                if grp > (N + bn - 1) // bn:
                    continue
                ops.set_tuning(bn, grp)
                sweep[f"bn{bn}_g{grp}"] = round(timed(lambda: ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True, out=y), 5) * 1e6, 1)
        ops.set_tuning(0, 0)
        res["sweep_us"] = sweep
        res["ok"] = True
    print("RESULT " + json.dumps(res))
    return 0 if res.get("ok") else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--filter", default=None)
    ap.add_argument("--out", default="gpurun_out/dev_lora_check.jsonl")
    args = ap.parse_args()
    if args.case:
        rc = 0
        for case in args.case.split(";"):
            try:
                rc |= run_case(case)
            except Exception as e:  # a CUDA error is sticky: stop this group
                print("RESULT " + json.dumps({"case": case, "ok": False, "exception": repr(e)[:800]}))
                rc = 1
                break
        sys.exit(rc)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    failures = 0
    with open(args.out, "w") as fout:
        kinds = []
        for case in CASES:
            k = case.split(":")[0]
            if k not in kinds:
                kinds.append(k)
        for kind in kinds:
            if args.filter and kind not in args.filter.split(","):
                continue
            group = [c for c in CASES if c.split(":")[0] == kind]
            t0 = time.time()
            try:
                pr = subprocess.run([sys.executable, __file__, "--case", ";".join(group)], capture_output=True, text=True,
                                    timeout=420)
                rc, out, err = pr.returncode, pr.stdout, pr.stderr
            except subprocess.TimeoutExpired as e:
                out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                rc, err = -999, "TIMEOUT"
            recs = [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
            seen = {r["case"] for r in recs}
            for c in group:
                if c not in seen:
                    recs.append({"case": c, "ok": False, "not_run_or_crashed": True})
            if rc != 0:
                recs.append({"case": f"<group {kind}>", "ok": False, "rc": rc, "stdout_tail": out[-1200:], "stderr_tail": err[-2500:]})
            for rec in recs:
                failures += 0 if rec.get("ok") else 1
                fout.write(json.dumps(rec) + "\n")
                print(json.dumps(rec)[:1500])
            fout.flush()
            print(f"[group {kind}] {time.time() - t0:.1f}s rc={rc}")
    print(f"failures: {failures}")


if __name__ == "__main__":
    main()
