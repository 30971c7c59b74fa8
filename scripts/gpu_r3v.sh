#!/bin/bash
# call 3v: pipeline-depth sweep of the headline bench (chains in flight)
OUT=gpurun_out
for P in 3 6 8 10 12; do
  timeout -k 10 400 python bench.py --pipeline $P --no-extras --no-cpu-baseline > $OUT/bench_r3v_p$P.json 2> $OUT/bench_r3v_p$P.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_r3v_p$P.json"))
print("P=$P", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]))
PY
done
