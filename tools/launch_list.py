"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.

    python tools/launch_list.py gpurun_out/x.csv [name-regex] [--second-half]
"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "."
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    recs = {}
    for r in rows[1:]:
        try:
            iid = int(r[ix["ID"]])
        except ValueError:
            continue
        d = recs.setdefault(iid, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
        d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    ids = [i for i in sorted(recs) if re.search(pat, recs[i]["name"])]
    if "--second-half" in sys.argv:
        ids = ids[len(ids) // 2:]
    tot = 0.0
    agg = {}
    for i in ids:
        d = recs[i]
        t, u = d["gpu__time_duration.sum"]
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[u]

        def mb(k):
            v, u = d.get(k, (0.0, "byte"))
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]

        rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
        tot += t
        nm = re.sub(r"\(.*", "", d["name"]).replace("void ", "")[:(60 if "--aggregate-only" in sys.argv else 30)]
        a = agg.setdefault(nm, [0.0, 0.0, 0.0, 0])
        a[0] += t; a[1] += rd; a[2] += wr; a[3] += 1
        if "--aggregate-only" not in sys.argv:
            print(f"{i:5d} {nm:30s} {d['grid']:>16s} {t:9.1f} us  rd {rd:7.0f} wr {wr:7.0f} MB  {(rd + wr) / t * 1e3:7.0f} GB/s")
    print(f"total {tot:.1f} us over {len(ids)} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:60s} {v[0]:9.1f} us  {v[0] / tot * 100:5.1f} %  rd {v[1]:7.0f} wr {v[2]:7.0f} MB  x{v[3]}")


if __name__ == "__main__":
    main()
