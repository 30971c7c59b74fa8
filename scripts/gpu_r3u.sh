#!/bin/bash
# call 3u: chain tests + bench with extras (other_configs with the SM-time F-FPS form)
OUT=gpurun_out
echo "== chain tests"; timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "chain" > $OUT/pytest_chain_r3u.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_chain_r3u.log | cut -c1-200
echo "== bench"; timeout -k 10 900 python bench.py --no-cpu-baseline > $OUT/bench_r3u.json 2> $OUT/bench_r3u.err; echo "bench rc=$?"; tail -2 $OUT/bench_r3u.err | cut -c1-300
python - <<PY
import json
d=json.load(open("$OUT/bench_r3u.json"))
print(d["value"], d["ms_per_step"])
for k,v in d["other_configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
PY
