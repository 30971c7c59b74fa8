#!/bin/bash
# Round 2 call W: D-FPS with the radix-sort prologue -- parity tests, timings, ncu source-level capture.
TAG=${1:-r2w}
OUT=gpurun_out
mkdir -p $OUT
echo "== fps tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "fps or live_reference or chain" > $OUT/pytest_fps_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_fps_$TAG.log | cut -c1-300
echo "== dfps time"; timeout -k 10 300 python scripts/dfps_time.py > $OUT/dfps_time_$TAG.log 2>&1; echo "rc=$?"; cat $OUT/dfps_time_$TAG.log | cut -c1-200
echo "== ncu dfps"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:fps_bucket" -s 1 -c 1 -f -o $OUT/prof_dfps_$TAG python scripts/ncu_fps.py > $OUT/prof_dfps_$TAG.log 2>&1; echo "rc=$?"; tail -2 $OUT/prof_dfps_$TAG.log | cut -c1-200
