#!/bin/bash
# call 3i: final validation of the round -- whole GPU suite, smoke, bench, ncu source-level capture of the 4-CTA F-FPS kernel.
TAG=${1:-r3i}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log | cut -c1-300
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log | cut -c1-300
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err | cut -c1-300; head -c 400 $OUT/bench_$TAG.json; echo
echo "== ncu ffps 4-CTA"; DE6D_PRUNE=1 DE6D_S=4 DE6D_BATCH=32 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:fps_features" -s 1 -c 1 -f -o $OUT/prof_ffps4_$TAG python scripts/ncu_ffps2.py > $OUT/prof_ffps4_$TAG.log 2>&1; echo "rc=$?"; tail -2 $OUT/prof_ffps4_$TAG.log | cut -c1-200
