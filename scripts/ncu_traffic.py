#!/usr/bin/env python
"""Builds profiles/ncu_traffic.json: DRAM bytes per C-ABI call (summed over the kernels the call launches) for the
bench workload, from an `ncu --set full` capture of scripts/ncu_chain.py (same chain as bench.py, batch 64, eager,
one stream) exported with `ncu -i X.ncu-rep --page raw --csv`.
    python scripts/ncu_traffic.py gpurun_out/prof_step_TAG_raw.csv
The capture is walked in lockstep with the entry-point sequence of one step (X.trace.json, written by ncu_chain.py
with DE6D_TRACE): each entry point expands to the kernels it launches."""
import csv
import json
import os
import sys

out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")


def kernels_of(entry, shape):
    if entry in ("de6d_furthest_point_sampling", "de6d_furthest_point_sampling_weights"):
        return ["fps_"]
    if entry == "de6d_furthest_point_sampling_features":
        return ["fps_features"]
    if entry == "de6d_furthest_point_sampling_matrix":
        return ["fps_matrix_kernel"]
    if entry == "de6d_dist_matrix":
        return ["dist_matrix_kernel"]
    if entry in ("de6d_gather_points", "de6d_group_points"):
        return ["group_"]
    if entry in ("de6d_group_concat", "de6d_group_concat_t"):
        return ["group_staged_kernel"]          # one launch: coordinate rows staged with the channels (r2)
    if entry == "de6d_gather_xyz":
        return ["gather_xyz_kernel"]
    if entry == "de6d_ball_query_grid_build":
        return ["bq_grid_build_kernel"]
    if entry == "de6d_ball_query_ex":
        if shape[1] == 3:
            return ["bq_grid_query_kernel"]     # grid built once per layer (de6d_ball_query_grid_build)
        return ["bq_grid_build_kernel", "bq_grid_query_kernel"] if shape[3] >= 2048 else ["ball_query_kernel"]
    if entry == "de6d_nms_batched":
        return ["nms_kernel"]
    return []


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
table = {}
for rep in sys.argv[1:]:
    trace = json.load(open(rep.replace("_raw.csv", ".trace.json")))
    rows = list(csv.reader(open(rep)))
    hdr, units = rows[0], rows[1]
    launches = [dict(zip(hdr, r)) for r in rows[2:]]

    def val(d, key):
        return float(d[key].replace(",", "")) * UNIT[units[hdr.index(key)]]

    pos = 0
    for entry, shape in trace:
        entry = entry.replace("_features_impl", "_features")
        names = kernels_of(entry, shape)
        rec = {"kernels": [], "dram_bytes": 0, "dram_read": 0, "dram_write": 0, "ncu_duration_us": 0.0, "report": os.path.basename(rep)}
        for want in names:
            assert pos < len(launches) and want in launches[pos]["Kernel Name"], (entry, shape, want, launches[pos]["Kernel Name"] if pos < len(launches) else None)
            d = launches[pos]
            pos += 1
            rec["kernels"].append(d["Kernel Name"].split("(")[0].replace("void ", ""))
            rec["dram_read"] += int(val(d, "dram__bytes_read.sum"))
            rec["dram_write"] += int(val(d, "dram__bytes_write.sum"))
            rec["ncu_duration_us"] += val(d, "gpu__time_duration.sum")
        rec["dram_bytes"] = rec["dram_read"] + rec["dram_write"]
        key = entry + ":" + ",".join(str(x) for x in shape)
        if key not in table:          # first occurrence of a shape (the three radius scales share sizes but not radii)
            table[key] = rec
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
print("wrote", out_path, len(table), "entries")
