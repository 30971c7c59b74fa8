#!/bin/bash
# call 3a: 4-CTA-cluster fused F-FPS (half of the channels from shared memory) -- parity tests, per-sample cost, batch sweep.
TAG=${1:-r3a}
OUT=gpurun_out
mkdir -p $OUT
echo "== ffps tests"; timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "ffps" > $OUT/pytest_ffps_$TAG.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_ffps_$TAG.log | cut -c1-300
echo "== m sweep"; timeout -k 10 300 python scripts/ffps_msweep.py > $OUT/ffps_msweep_$TAG.log 2>&1; echo "rc=$?"; cat $OUT/ffps_msweep_$TAG.log | cut -c1-250
echo "== variants"; timeout -k 10 600 python scripts/ffps_variants.py $OUT/ffps_variants_$TAG.json > $OUT/ffps_variants_$TAG.log 2>&1; echo "rc=$?"; tail -40 $OUT/ffps_variants_$TAG.log | cut -c1-200
