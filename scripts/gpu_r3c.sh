#!/bin/bash
# call 3c: the whole GPU suite, smoke, bench (+ reference arm), ncu launch list + full capture of one chain step.
TAG=${1:-r3c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log | cut -c1-300
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err | cut -c1-300; head -c 400 $OUT/bench_$TAG.json; echo
echo "== bench reference arm"; timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "rc=$?"; head -c 300 $OUT/bench_ref_$TAG.json; echo
bash scripts/gpu_prof_r2.sh $TAG
