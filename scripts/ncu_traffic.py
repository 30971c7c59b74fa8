#!/usr/bin/env python
"""Builds profiles/ncu_traffic.json: DRAM bytes per launch of the kernels behind the bench's C-ABI entry points, read
from `ncu --set full` reports of scripts/ncu_chain.py (same workload as bench.py, batch 64, eager, one stream).
    python scripts/ncu_traffic.py gpurun_out/prof_A.ncu-rep [more.ncu-rep ...]
The mapping kernel -> (entry point, shape) is positional: ncu_chain.py prints the entry-point sequence of one step
(DE6D_TRACE=1) into <rep>.trace.json, and the k-th captured launch of a kernel name is matched to the k-th call whose
entry point launches that kernel."""
import csv
import io
import json
import os
import subprocess
import sys

KERNEL_OF = {   # kernel-name substring -> entry point
    "fps_bucket_kernel<0": "de6d_furthest_point_sampling", "fps_bucket_kernel<1": "de6d_furthest_point_sampling_weights",
    "fps_features_kernel": "de6d_furthest_point_sampling_features", "group_staged_kernel": "de6d_group_concat",
    "group_direct_kernel": "de6d_group_concat",
    "bq_grid_query_kernel": "de6d_ball_query_ex", "nms_kernel": "de6d_nms_batched", "dist_matrix_kernel": "de6d_dist_matrix",
    "fps_matrix_kernel": "de6d_furthest_point_sampling_matrix",
}
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
table = json.load(open(out_path)) if os.path.exists(out_path) else {}
for rep in sys.argv[1:]:
    trace = json.load(open(rep.replace(".ncu-rep", ".trace.json")))   # [[entry, [shape...]], ...] of ONE step
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    seen = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        entry = next((e for k, e in KERNEL_OF.items() if k in name), None)
        if entry is None:
            continue
        calls = [t for t in trace if t[0] == entry]
        i = seen.get(entry, 0)
        seen[entry] = i + 1
        if not calls:
            continue
        shape = calls[i % len(calls)][1]
        def val(key):
            v, u = float(d[key].replace(",", "")), rows[1][hdr.index(key)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        key = entry + ":" + ",".join(str(x) for x in shape)
        table[key] = {"kernel": name.split("(")[0], "dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                      "dram_read": int(val("dram__bytes_read.sum")), "dram_write": int(val("dram__bytes_write.sum")),
                      "ncu_duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}[rows[1][hdr.index("gpu__time_duration.sum")]],
                      "report": os.path.basename(rep)}
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
print("wrote", out_path, len(table), "entries")
