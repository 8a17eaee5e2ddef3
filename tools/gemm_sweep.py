"""Device-only timing of the fused projection + LoRA kernel per BASELINE shape under different tile / group choices.

    python tools/gemm_sweep.py [--plain] [--bwd] [--shapes "M,K,N,tok;..."] [--bn 0,128,160,192] [--group 0,1,2,4]

Each configuration is captured as a CUDA graph of `copies` launches over DISTINCT operand sets (sized to exceed the 126 MB L2) and
timed with CUDA events around back-to-back replays -- the same method as bench.py's roofline probe."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aqualora_b200 import ops

SHAPES = [(65536, 320, 320, 4096), (16384, 640, 640, 1024), (4096, 1280, 1280, 256), (65536, 320, 2560, 4096), (65536, 1280, 320, 4096),
          (16384, 640, 5120, 1024), (16384, 2560, 640, 1024), (4096, 1280, 10240, 256), (4096, 5120, 1280, 256)]


def time_graph(fn, min_ms=150.0):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    for _ in range(2):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    reps = int(min(200, max(3, min_ms / max(e0.elapsed_time(e1), 1e-3))))
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plain", action="store_true")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--shapes", default="")
    ap.add_argument("--bn", default="0")
    ap.add_argument("--group", default="0")
    ap.add_argument("--rank", type=int, default=64)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    shapes = SHAPES if not a.shapes else [tuple(int(v) for v in it.split(",")) for it in a.shapes.split(";")]
    bns = [int(v) for v in a.bn.split(",")]
    groups = [int(v) for v in a.group.split(",")]
    r = a.rank
    for (M, K, N, tok) in shapes:
        per = 2 * M * (K + N) + 2 * M * r
        copies = max(2, min(16, int(3e8 // per) + 1))
        gen = torch.Generator(device=dev).manual_seed(0)
        xs = [torch.randn(M, K, generator=gen, device=dev).bfloat16() for _ in range(copies)]
        w = (torch.randn(N, K, generator=gen, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(N, generator=gen, device=dev).bfloat16()
        dn = (torch.randn(r, K, generator=gen, device=dev) * K ** -0.5).bfloat16()
        up = (torch.randn(N, r, generator=gen, device=dev) * 0.1).bfloat16()
        sc = torch.randn(M // tok, r, generator=gen, device=dev)
        flops = 2.0 * M * K * N + (0 if a.plain else 2.0 * M * r * (K + N))
        if a.bwd:
            gys = [torch.randn(M, N, generator=gen, device=dev).bfloat16() for _ in range(copies)]
            hs = [ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True)[1] for x in xs]
            wt, dnt, upt = w.t().contiguous(), dn.t().contiguous(), up.t().contiguous()
            g_dn, g_up, g_sc = torch.zeros(r, K, device=dev), torch.zeros(N, r, device=dev), torch.zeros(M // tok, r, device=dev)
            flops = 2.0 * M * K * N + 4.0 * M * r * (K + N)
        for bn in bns:
            for grp in groups:
                ops.set_tuning(bn, grp)
                try:
                    if a.bwd:
                        fn = lambda: [ops.lora_linear_bwd(gy, x, wt, dnt, upt, sc, h, g_dn, g_up, g_sc, tok) for gy, x, h in zip(gys, xs, hs)]
                    elif a.plain:
                        fn = lambda: [ops.lora_linear_fwd(x, w, b, None, None, None, tok) for x in xs]
                    else:
                        fn = lambda: [ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True) for x in xs]
                    ms = time_graph(fn) / copies
                    print(f"M={M:6d} K={K:5d} N={N:6d} bn={bn:3d} group={grp:2d}  {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.0f} TFLOP/s", flush=True)
                except Exception as e:
                    print(f"M={M:6d} K={K:5d} N={N:6d} bn={bn:3d} group={grp:2d}  failed: {str(e)[:80]}", flush=True)
                finally:
                    ops.set_tuning(0, 0)
        del xs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
