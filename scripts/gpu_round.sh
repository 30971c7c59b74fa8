#!/bin/bash
# One gpurun call: golden vectors from the reference CUDA kernels, GPU parity suite, smoke, bench, ncu launch list.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== golden" ; timeout 600 python tests/golden/make_golden.py --cuda --out $OUT/golden_cuda.npz > $OUT/golden_$TAG.log 2>&1 ; echo "golden rc=$?"
cp -f $OUT/golden_cuda.npz tests/golden/golden_cuda.npz 2>/dev/null
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1 ; echo "pytest rc=$?" ; tail -30 $OUT/pytest_gpu_$TAG.log
echo "== pytest cpu(oracle vs golden_cuda)" ; timeout 600 python -m pytest tests/test_oracle.py -q > $OUT/pytest_oracle_$TAG.log 2>&1 ; echo "rc=$?" ; tail -5 $OUT/pytest_oracle_$TAG.log
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1 ; echo "smoke rc=$?" ; tail -5 $OUT/smoke_$TAG.log
echo "== bench" ; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ; echo "bench rc=$?" ; tail -3 $OUT/bench_$TAG.err ; head -c 3000 $OUT/bench_$TAG.json
echo "== ncu launches" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv python scripts/ncu_chain.py > $OUT/ncu_launches_$TAG.log 2>&1 ; echo "ncu rc=$?"
