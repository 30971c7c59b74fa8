#!/bin/bash
# Round 2, call A: parity suite incl. the drop-in tests (unmodified reference python over compat), fused F-FPS vs the
# reference pipeline, compute-sanitizer on the small-cloud FPS kernel, per-op table incl. the 131072-point stress shapes.
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/gpu_${TAG}.txt 2>&1
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_${TAG}.log
echo "== ffps agreement"; timeout -k 10 600 python scripts/ffps_agreement.py --out $OUT/r2_ffps_agreement_${TAG}.json > $OUT/ffps_agreement_${TAG}.log 2>&1; echo "rc=$?"; tail -3 $OUT/ffps_agreement_${TAG}.log | cut -c1-600
echo "== sanitizer fps_small"
for tool in memcheck racecheck synccheck; do
  timeout -k 10 700 compute-sanitizer --tool $tool --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_$tool.log \
     python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 600 -k "small_cloud" -p no:cacheprovider > $OUT/sanitizer_${TAG}_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 $OUT/sanitizer_${TAG}_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|errors" $OUT/sanitizer_${TAG}_$tool.log | tail -3
done
echo "== ops"; timeout -k 10 900 python scripts/bench_ops.py > $OUT/ops_${TAG}.txt 2>&1; echo "rc=$?"; tail -12 $OUT/ops_${TAG}.txt
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -3 $OUT/bench_${TAG}.err; head -c 600 $OUT/bench_${TAG}.json; echo
