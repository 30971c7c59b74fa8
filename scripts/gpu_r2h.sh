#!/bin/bash
# Round 2, call H: fused-SA tests, bench with the sa_mlp leg, ncu evidence (launch list, full captures).
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
echo "== tests"; timeout -k 10 600 python -m pytest tests/test_sa_fused_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 300 -k "sa_ or fused or cluster_sizes or wrapper_golden or full_pose" > $OUT/pytest_sel_${TAG}.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_sel_${TAG}.log | cut -c1-300
echo "== bench"; timeout -k 10 1200 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -5 $OUT/bench_${TAG}.err; head -c 300 $OUT/bench_${TAG}.json; echo
bash scripts/gpu_prof_r2.sh ${TAG}
