#!/bin/bash
# call 3d: compute-sanitizer on the kernels added / changed late in round 2 (pruned and 4-CTA F-FPS, D-FPS radix prologue + paired
# visits, fused SA scale with the last layer split over launches).
TAG=${1:-r3d}
OUT=gpurun_out
mkdir -p $OUT
SEL='pruned_equals_dense or pruned_non_finite or cluster_sizes_agree or multi_sample_rounds or ffps_adversarial'
for tool in memcheck synccheck; do
  timeout -k 10 1200 compute-sanitizer --tool $tool --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_$tool.log \
     python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 1000 -k "$SEL" -p no:cacheprovider > $OUT/sanitizer_${TAG}_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 $OUT/sanitizer_${TAG}_${tool}_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_${TAG}_$tool.log | tail -2
done
RSEL='(pruned_equals_dense and (384 or 1000 or 65-3 or 2048)) or (cluster_sizes_agree and 3600) or (multi_sample_rounds and 3000)'
timeout -k 10 1500 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_racecheck.log \
   python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 1400 -k "$RSEL" -p no:cacheprovider > $OUT/sanitizer_${TAG}_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 $OUT/sanitizer_${TAG}_racecheck_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_${TAG}_racecheck.log | tail -2
echo "== sa_mlp slice under memcheck"
timeout -k 10 600 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_samlp_memcheck.log \
   python -m pytest tests/test_sa_fused_gpu.py -m gpu -q -x --timeout 500 -k "131-128-256-256 or 67-128-256-256 or rejects" -p no:cacheprovider > $OUT/sanitizer_${TAG}_samlp_pytest.log 2>&1
echo "sa_mlp memcheck rc=$?"; tail -2 $OUT/sanitizer_${TAG}_samlp_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY" $OUT/sanitizer_${TAG}_samlp_memcheck.log | tail -2
