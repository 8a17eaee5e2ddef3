"""Developer tool: per-role timeline of the fused projection + LoRA kernel on one shape.

    python tools/gemm_trace.py build                      # here (no GPU): tools/_trace/libaq_trace.so with -DAQ_GEMM_TRACE
    AQUALORA_B200_LIB=tools/_trace/libaq_trace.so python tools/gemm_trace.py run "65536,320,320,4096;4096,1280,1280,256"

CTA 0 stamps clock64() at the hand-over points of its producer / MMA / epilogue warps (slots in csrc/lora_gemm.cu); the run
prints, per work item, every stamp relative to the kernel's first stamp in microseconds.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_trace")
SLOTS = ["prod_begin", "prod_end", "mma_acc0", "mma_first_full", "mma_H_commit", "mma_acc1", "mma_t1_issued", "mma_hs_ready",
         "mma_acc0_commit", "mma_acc1_commit", "epi_begin", "epi_H_full", "epi_hs_arrive", "epi_t0_full", "epi_t0_drained",
         "epi_t0_stored", "epi_t1_full", "epi_t1_drained", "epi_t1_stored", "epi_t2_full", "epi_t2_drained", "epi_t2_stored",
         "epi_t3_full", "epi_t3_drained", "epi_t3_stored", "t0_p0_staged", "t0_p0_readback", "t0_p0_stg_issued", "t0_p1_staged",
         "t0_p1_readback", "t0_p1_stg_issued"]


def build():
    from aqualora_b200 import build as B
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for src in B.sources():
        obj = os.path.join(OUT, src.stem + ".o")
        cmd = [B._nvcc(), *B.NVCC_FLAGS, "-DAQ_GEMM_TRACE", "-c", str(src), "-o", obj]
        subprocess.run(cmd, check=True, capture_output=True)
        objs.append(obj)
    subprocess.run([B._nvcc(), "-shared", "-o", os.path.join(OUT, "libaq_trace.so"), *objs, "-lcudart"], check=True)
    print(os.path.join(OUT, "libaq_trace.so"))


def run(spec, plain=False):
    import torch
    from aqualora_b200 import _lib, ops
    lib = _lib.load()
    lib.aq_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.aq_debug_gemm_cta_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    dev = torch.device("cuda:0")
    mhz = 1965.0   # SM clock under load on this pool (bench.py's clocks line)
    for item in spec.split(";"):
        M, K, N, tok = (int(v) for v in item.split(","))
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(M, K, generator=g, device=dev).bfloat16()
        w = (torch.randn(N, K, generator=g, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(N, generator=g, device=dev).bfloat16()
        dn = (torch.randn(64, K, generator=g, device=dev) * K ** -0.5).bfloat16()
        up = (torch.randn(N, 64, generator=g, device=dev) * 0.1).bfloat16()
        sc = torch.randn(M // tok, 64, generator=g, device=dev)
        for _ in range(3):
            ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"== M={M} K={K} N={N}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch back-to-back (incl. trace stores)")
        buf = (ctypes.c_ulonglong * (64 * 32))()
        ct = (ctypes.c_ulonglong * (256 * 4))()
        lib.aq_debug_gemm_cta_times(ct, 1)            # clear
        torch.cuda.synchronize()
        ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True)
        ops.lora_linear_fwd(x, w, b, dn, up, sc, tok, save_h=True) if os.environ.get("AQ_TRACE_TWICE") else None
        lib.aq_debug_gemm_cta_times(ct, 0)
        rows = [tuple(ct[i * 4: i * 4 + 4]) for i in range(256) if ct[i * 4]]
        if rows:
            s0 = min(r[0] for r in rows)
            starts = sorted(r[0] - s0 for r in rows)
            pro = sorted(r[1] - r[0] for r in rows)
            ends = sorted(r[2] - s0 for r in rows)
            print(f" {len(rows)} CTAs on {len(set(r[3] for r in rows))} SMs: entry spread {starts[-1] / 1e3:.2f} us, prologue min/med/max "
                  f"{pro[0] / 1e3:.2f}/{pro[len(pro) // 2] / 1e3:.2f}/{pro[-1] / 1e3:.2f} us, exit min/med/max "
                  f"{ends[0] / 1e3:.2f}/{ends[len(ends) // 2] / 1e3:.2f}/{ends[-1] / 1e3:.2f} us after the first entry; CTA0 "
                  f"entry {(rows[0][0] - s0) / 1e3:.2f} prologue {(rows[0][1] - rows[0][0]) / 1e3:.2f} exit {(rows[0][2] - s0) / 1e3:.2f}")
        rc = lib.aq_debug_gemm_trace(buf, 64 * 32)
        assert rc == 0, rc
        vals = list(buf)
        t0 = min(v for v in vals[:32] if v)
        for it in range(8):
            row = vals[it * 32: it * 32 + 31]
            if not any(row):
                break
            print(f" item {it}: " + "  ".join(f"{n}={(v - t0) / mhz:.2f}" for n, v in zip(SLOTS, row) if v))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2])
