#!/usr/bin/env python
"""Builds profiles/ncu_issue.json: issue-slot statistics of the latency/issue-bound sampling kernels of the bench workload from
an `ncu --set full` capture of scripts/ncu_chain.py exported with `ncu -i X.ncu-rep --page raw --csv` (same capture as
scripts/ncu_traffic.py):
    python scripts/ncu_issue.py gpurun_out/prof_step_TAG_raw.csv
issue_active      smsp__issue_active.avg.pct_of_peak_sustained_active / 100 (over the SMs that ran the kernel)
sms_occupied      SMs holding at least one CTA = sm__cycles_active.sum / sm__cycles_active.max, capped at 148
cycles_per_sample sm__cycles_active.max / npoint (samples drawn per cloud)
smem_wavefront_frac  shared-memory LSU wavefronts / (sms_occupied * active cycles): 1 wavefront per cycle per SM is the port's peak
stalls            smsp__average_warps_issue_stalled_*_per_issue_active ratios, normalised to their sum (top five)"""
import csv
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(root, "profiles", "ncu_issue.json")
UNIT = {"": 1.0, "cycle": 1.0, "%": 1.0}
table = {}
for rep in sys.argv[1:]:
    trace = json.load(open(rep.replace("_raw.csv", ".trace.json")))
    rows = list(csv.reader(open(rep)))
    hdr = rows[0]
    launches = [dict(zip(hdr, r)) for r in rows[2:]]

    def val(d, key):
        v = d.get(key, "")
        return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0

    want = {"de6d_furthest_point_sampling": "fps_", "de6d_furthest_point_sampling_weights": "fps_",
            "de6d_furthest_point_sampling_features": "fps_features", "de6d_furthest_point_sampling_matrix": "fps_matrix_kernel"}
    fps_launches = [d for d in launches if "fps_" in d["Kernel Name"]]
    pos = 0
    for entry, shape in trace:
        entry = entry.replace("_features_impl", "_features")
        if entry not in want:
            continue
        d = fps_launches[pos]
        pos += 1
        assert want[entry] in d["Kernel Name"], (entry, d["Kernel Name"])
        m = shape[3] if entry == "de6d_furthest_point_sampling_features" else shape[2]
        act_max, act_sum = val(d, "sm__cycles_active.max"), val(d, "sm__cycles_active.sum")
        sms = min(148.0, act_sum / act_max) if act_max else 0.0
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: val(d, k) for k in hdr
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k}
        tot = sum(stalls.values()) or 1.0
        top = dict(sorted(((k, round(v / tot, 3)) for k, v in stalls.items()), key=lambda kv: -kv[1])[:5])
        wave = val(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        key = entry + ":" + ",".join(str(x) for x in shape)
        if key in table:
            continue
        table[key] = {"kernel": d["Kernel Name"].split("(")[0].replace("void ", ""),
                      "issue_active": round(val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0, 4),
                      "sms_occupied": round(sms, 1), "cycles_per_sample": round(act_max / max(m, 1), 1),
                      "smem_wavefront_frac": round(wave / act_sum, 4) if act_sum else None, "stalls": top,
                      "grid": int(val(d, "launch__grid_size")), "cluster": int(val(d, "launch__cluster_size") or 1),
                      "ncu_duration_us": round(val(d, "gpu__time_duration.sum"), 1), "source": "profiles/" + os.path.basename(rep)}
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
print("wrote", out_path, len(table), "entries")
for k, v in table.items():
    print(k, v["issue_active"], v["sms_occupied"], v["cycles_per_sample"], v["smem_wavefront_frac"], v["stalls"])
