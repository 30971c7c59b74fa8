#!/bin/bash
# multi-GPU bench exactly as the driver launches it: torchrun, one rank per GPU.  Usage: bash scripts/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_${TAG}_${N}gpu.json 2> $OUT/bench_${TAG}_${N}gpu.err; echo "rc=$?"; tail -3 $OUT/bench_${TAG}_${N}gpu.err | cut -c1-300; python - <<PY
import json
d=json.load(open("$OUT/bench_${TAG}_${N}gpu.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "e2e_all", d["e2e_all_inputs"]["value"], "gather ms", d["detections_all_gather_ms"], d["clocks"])
PY
