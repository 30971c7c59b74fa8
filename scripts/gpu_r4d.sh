#!/bin/bash
# call 4d: FPS tests (incl. cluster D-FPS at 131072 points, S-FPS, one-sample path) + bench with all legs
OUT=gpurun_out
echo "== fps tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "fps or live_reference or chain" > $OUT/pytest_fps_r4d.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_fps_r4d.log | cut -c1-200
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/bench_r4d.json 2> $OUT/bench_r4d.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r4d.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "fps_us", d["fps_us_per_frame"], "serialized", d["serialized_kernel_ms_per_step"])
for k,v in d["other_configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("serial_ms_per_step"))
PY
