#!/bin/bash
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
echo "== tests (pair shapes first)"; timeout -k 10 240 python -m pytest tests/test_sa_fused_gpu.py -m gpu -q --timeout 120 -x -k "131" > $OUT/pytest_sapair_${TAG}.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_sapair_${TAG}.log | cut -c1-300
echo "== tests (all sa)"; timeout -k 10 400 python -m pytest tests/test_sa_fused_gpu.py -m gpu -q --timeout 200 > $OUT/pytest_sa_${TAG}.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_sa_${TAG}.log | cut -c1-300
nvidia-smi --query-gpu=name,memory.used --format=csv
