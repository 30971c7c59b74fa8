#!/bin/bash
# One gpurun call: (golden vectors from the reference CUDA kernels,) GPU parity suite, smoke, bench, ncu launch list and
# a full ncu capture of one chain step.  Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [golden]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/gpu_$TAG.txt 2>&1
if [ "$2" = "golden" ]; then
  echo "== golden" ; timeout 600 python tests/golden/make_golden.py --cuda --out $OUT/golden_cuda.npz > $OUT/golden_$TAG.log 2>&1 ; echo "golden rc=$?"
fi
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1 ; echo "pytest rc=$?" ; tail -4 $OUT/pytest_gpu_$TAG.log
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1 ; echo "smoke rc=$?" ; tail -2 $OUT/smoke_$TAG.log
echo "== bench" ; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; echo "bench rc=$?" ; tail -3 $OUT/bench_$TAG.err ; head -c 400 $OUT/bench_$TAG.json; echo
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err ; echo "rc=$?"; head -c 300 $OUT/bench_ref_$TAG.json; echo
echo "== ncu launches" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv python scripts/ncu_chain.py > $OUT/ncu_launches_$TAG.log 2>&1 ; echo "ncu rc=$?"
echo "== ncu full, one step" ; DE6D_STEPS=0 DE6D_TRACE=$OUT/prof_step_$TAG.trace.json timeout 1500 ncu --set full --clock-control none -k "regex:fps_|group_|bq_grid|nms_kernel|ball_query_kernel|dist_matrix|iou_matrix" -c 80 -f -o /tmp/prof_step_$TAG python scripts/ncu_chain.py > $OUT/prof_step_$TAG.log 2>&1 ; echo "ncu rc=$?"; ncu -i /tmp/prof_step_$TAG.ncu-rep --page raw --csv > $OUT/prof_step_${TAG}_raw.csv 2>/dev/null; ls -la $OUT/prof_step_${TAG}_raw.csv
