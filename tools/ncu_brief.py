"""Headline metrics + issue-stall ratios of every launch in .ncu-rep files.    python tools/ncu_brief.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size"]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, units = rows[0], rows[1]
        ix = {k: i for i, k in enumerate(h)}
        for d in rows[2:]:
            print(f"=== {rep}: {d[ix['Kernel Name']][:100]}")
            for w in WANT:
                if w in ix:
                    print(f"   {w:78s} {d[ix[w]]} {units[ix[w]]}")
            for k in h:
                if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                    try:
                        v = float(d[ix[k]])
                    except ValueError:
                        continue
                    if v > 0.3:
                        print(f"   stall {k[34:-23]:60s} {v:.2f} warps per issue")


if __name__ == "__main__":
    main()
