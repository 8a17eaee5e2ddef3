"""Stall samples of the fused GEMM kernel per warp role (regions split at the USETMAXREG instructions) from an .ncu-rep
captured with --import-source on.    python tools/ncu_src_regions.py x.ncu-rep [top-n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    jx = {k: i for i, k in enumerate(h)}
    body, seen = [], set()
    for r in rows[2:]:
        if not r or not r[0].startswith("0x"):
            continue
        if r[0] in seen:
            break
        seen.add(r[0])
        body.append(r)
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k and "not_issued" not in k.lower()]
    cuts = [i for i, r in enumerate(body) if "USETMAXREG" in r[jx["Source"]]]
    names = ["prologue"] + [f"region after USETMAXREG #{k} ({body[c][jx['Source']].strip()[:40]})" for k, c in enumerate(cuts)]
    bounds = [0] + cuts + [len(body)]
    for n in range(len(bounds) - 1):
        seg = body[bounds[n]:bounds[n + 1]]
        tot = sum(int(r[jx["# Samples"]]) for r in seg)
        ex = sum(int(r[jx["Instructions Executed"]]) for r in seg)
        agg = {}
        for k in stalls:
            v = sum(int(r[jx[k]]) for r in seg)
            if v:
                agg[k[6:]] = v
        print(f"== {names[n]}: {len(seg)} instr, {tot} samples, {ex} warp-instr executed")
        print("   " + ", ".join(f"{k}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
        top = sorted(range(len(seg)), key=lambda i: -int(seg[i][jx["# Samples"]]))[:topn]
        for i in sorted(top):
            r = seg[i]
            s = {k[6:]: int(r[jx[k]]) for k in stalls if int(r[jx[k]]) > 0}
            print(f"   {bounds[n] + i:5d} {r[jx['# Samples']]:>6s}  {r[jx['Source']].strip()[:64]:64s} {s}")


if __name__ == "__main__":
    main()
