#!/bin/bash
# Round 2 call T: pruned F-FPS kernel -- parity tests, then dense vs pruned timings.
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
echo "== ffps tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "ffps" > $OUT/pytest_ffps_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_ffps_$TAG.log | cut -c1-300
echo "== variants"; timeout -k 10 600 python scripts/ffps_variants.py $OUT/ffps_variants_$TAG.json > $OUT/ffps_variants_$TAG.log 2>&1; echo "rc=$?"; tail -62 $OUT/ffps_variants_$TAG.log | cut -c1-200
