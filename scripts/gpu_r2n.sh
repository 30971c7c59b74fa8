#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out
mkdir -p $OUT
echo "== sa leg"; timeout -k 10 600 python -c "
import json, torch, bench
print(json.dumps(bench.sa_mlp_leg(64, torch.device('cuda', 0)), indent=1))" > $OUT/sa_leg_${TAG}.json 2> $OUT/sa_leg_${TAG}.err; echo "rc=$?"; cat $OUT/sa_leg_${TAG}.json | head -50; tail -3 $OUT/sa_leg_${TAG}.err
echo "== pytest gpu (all)"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_${TAG}.log | cut -c1-300
