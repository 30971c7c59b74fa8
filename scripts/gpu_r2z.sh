#!/bin/bash
# Round 2 call Z: fused SA scale with the last layer split over launches -- tests; full GPU suite; bench line.
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
echo "== sa fused tests"; timeout -k 10 900 python -m pytest tests/test_sa_fused_gpu.py -m gpu -q -x --timeout 600 > $OUT/pytest_sa_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_sa_$TAG.log | cut -c1-300
echo "== full gpu suite"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log | cut -c1-300
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err | cut -c1-300; head -c 600 $OUT/bench_$TAG.json; echo
