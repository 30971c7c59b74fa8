#!/bin/bash
# call 3l: whole GPU suite + smoke on the final library (new F-FPS layout / batch tests included).
TAG=${1:-r3l}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log | cut -c1-300
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log | cut -c1-300
