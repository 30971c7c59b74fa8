#!/bin/bash
# call 3f: ball query with four candidates in flight -- parity tests, then the bench's per-kernel pass.
TAG=${1:-r3f}
OUT=gpurun_out
mkdir -p $OUT
echo "== ball query / chain / dropin tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin_gpu.py -m gpu -q -x --timeout 600 -k "ball or query or chain or shared_grid or live_reference or golden or sa_module or backbone or ffps" > $OUT/pytest_bq_$TAG.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_bq_$TAG.log | cut -c1-300
echo "== bench (no extras)"; timeout -k 10 600 python bench.py --no-extras --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err | cut -c1-300; head -c 300 $OUT/bench_$TAG.json; echo
