#!/bin/bash
# Round 2 call X: D-FPS with 6 / 8 samples per round -- exactness tests + timings.
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
echo "== fps tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "fps or live_reference" > $OUT/pytest_fps_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_fps_$TAG.log | cut -c1-300
echo "== dfps time"; timeout -k 10 300 python scripts/dfps_time.py > $OUT/dfps_time_$TAG.log 2>&1; echo "rc=$?"; cat $OUT/dfps_time_$TAG.log | cut -c1-200
