#!/bin/bash
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
echo "== tests"; timeout -k 10 600 python -m pytest tests/test_sa_fused_gpu.py -m gpu -q --timeout 300 > $OUT/pytest_sa_${TAG}.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_sa_${TAG}.log | cut -c1-300
echo "== sa leg"; timeout -k 10 600 python -c "
import json, torch, bench
print(json.dumps(bench.sa_mlp_leg(64, torch.device('cuda', 0)), indent=1))" > $OUT/sa_leg_${TAG}.json 2> $OUT/sa_leg_${TAG}.err; echo "rc=$?"; cat $OUT/sa_leg_${TAG}.json | head -30; tail -3 $OUT/sa_leg_${TAG}.err
echo "== ncu full, fused SA"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:sa_mlp_kernel" -c 6 -f -o /tmp/prof_samlp_$TAG python scripts/ncu_sa_mlp.py > $OUT/prof_samlp_$TAG.log 2>&1; echo "ncu rc=$?"; ncu -i /tmp/prof_samlp_$TAG.ncu-rep --page raw --csv > $OUT/prof_samlp_${TAG}_raw.csv 2>/dev/null; ncu -i /tmp/prof_samlp_$TAG.ncu-rep --page source --csv > $OUT/prof_samlp_${TAG}_source.csv 2>/dev/null; ls -la $OUT/prof_samlp_${TAG}_*.csv
