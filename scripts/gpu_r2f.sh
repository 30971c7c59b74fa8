#!/bin/bash
# Round 2, call F: fused SA-MLP (tcgen05) tests first (bounded), then the whole suite.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
echo "== sa_fused"; timeout -k 10 300 python -m pytest tests/test_sa_fused_gpu.py -m gpu -q --timeout 200 > $OUT/pytest_sa_${TAG}.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_sa_${TAG}.log | cut -c1-300
echo "== golden2"; timeout -k 10 300 python tests/golden/make_golden.py --cuda2 > $OUT/golden2_${TAG}.log 2>&1; echo "rc=$?"; tail -2 $OUT/golden2_${TAG}.log
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_sa_fused_gpu.py > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu_${TAG}.log | cut -c1-300
