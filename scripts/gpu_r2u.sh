#!/bin/bash
# Round 2 call U: D-FPS with paired bucket visits -- parity tests + timings; F-FPS m-sweep (prologue vs per-sample cost).
TAG=${1:-r2u}
OUT=gpurun_out
mkdir -p $OUT
echo "== fps tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "fps or live_reference or chain" > $OUT/pytest_fps_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_fps_$TAG.log | cut -c1-300
echo "== dfps time"; timeout -k 10 300 python scripts/dfps_time.py > $OUT/dfps_time_$TAG.log 2>&1; echo "rc=$?"; cat $OUT/dfps_time_$TAG.log | cut -c1-200
echo "== ffps m sweep"; timeout -k 10 300 python scripts/ffps_msweep.py > $OUT/ffps_msweep_$TAG.log 2>&1; echo "rc=$?"; cat $OUT/ffps_msweep_$TAG.log | cut -c1-250
