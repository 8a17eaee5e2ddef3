"""Summarise an .ncu-rep offline: per-launch headline metrics and (optionally) the top stall sites of one launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--src KERNEL_REGEX:INDEX]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, data = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    print("launches:", len(data))
    for n, d in enumerate(data):
        print(f"--- [{n}] {d[ix['Kernel Name']][:90]}  grid {d[ix.get('Grid Size', 0)]}")
        for m in METRICS:
            if m in ix:
                print(f"    {m:72s} {d[ix[m]]:>16s} {units[ix[m]]}")
    if "--src" in sys.argv:
        sel = sys.argv[sys.argv.index("--src") + 1]
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{sel}"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        body = [r for r in rows[2:] if r and r[0].startswith("0x")]
        # the source page lists the function once per view; keep the first copy
        seen, uniq = set(), []
        for r in body:
            if r[0] in seen:
                break
            seen.add(r[0])
            uniq.append(r)
        jx = {k: i for i, k in enumerate(h)}
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        tot = sum(int(r[jx["# Samples"]]) for r in uniq)
        print(f"source view: {len(uniq)} instructions, {tot} samples")
        top = sorted(range(len(uniq)), key=lambda i: -int(uniq[i][jx["# Samples"]]))[:int(sys.argv[sys.argv.index("--src") + 2]) if len(sys.argv) > sys.argv.index("--src") + 2 else 40]
        for i in sorted(top):
            r = uniq[i]
            s = {k[6:]: int(r[jx[k]]) for k in stalls if int(r[jx[k]]) > 0}
            print(f"  {i:5d} {r[jx['# Samples']]:>6s}  {r[jx['Source']].strip()[:70]:70s} {s}")


if __name__ == "__main__":
    main()
