#!/usr/bin/env python
"""Per-kernel headline metrics from an `ncu --page raw --csv` export (one row per launch): kernels grouped by name.
    python scripts/ncu_raw_summary.py profiles/X_raw.csv > profiles/X_summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0,
        "nsecond": 1e-3}


def val(r, name, scale=False):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return 0.0
    v = float(r[i].replace(",", ""))
    return v * UNIT.get(units[i].split("/")[0], 1.0) if scale else v


agg = collections.OrderedDict()
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("de6d::", "")
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, issue=0.0, warps=0.0, fma=0.0, regs=0, smem=0.0, grid=0, block=0))
    a["n"] += 1
    a["us"] += val(r, "gpu__time_duration.sum", True)
    a["rd"] += val(r, "dram__bytes_read.sum", True)
    a["wr"] += val(r, "dram__bytes_write.sum", True)
    a["issue"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["warps"] += val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["fma"] += val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
    a["regs"] = int(val(r, "launch__registers_per_thread"))
    a["smem"] = val(r, "launch__shared_mem_per_block_dynamic", True) + val(r, "launch__shared_mem_per_block_static", True)
    a["grid"] = int(val(r, "launch__grid_size")); a["block"] = int(val(r, "launch__block_size"))
tot = sum(a["us"] for a in agg.values())
print("chain steps (n = launches captured per kernel; the capture spans the warm-up step and the traced step), batch 64, eager, single stream, `ncu --set full --clock-control none` (cold-cache, serialised: use the SHARES)")
print("%-44s %3s %9s %6s %9s %9s %7s %7s %6s %5s %8s %6s" % ("kernel", "n", "us", "share", "dram rd MB", "dram wr MB", "GB/s", "issue%", "warps%", "regs", "smem KB", "block"))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
    print("%-44s %3d %9.1f %5.1f%% %9.1f %9.1f %7.0f %7.1f %6.1f %5d %8.1f %6d" % (
        name[:44], n, a["us"], 100 * a["us"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, gbs, a["issue"] / n, a["warps"] / n,
        a["regs"], a["smem"] / 1e3, a["block"]))
print("total %.1f us" % tot)
