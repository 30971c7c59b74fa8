#!/bin/bash
# call 3y: synccheck (+ memcheck) after the explicit reconvergence behind the mbarrier spin loops; then the plain suite.
OUT=gpurun_out
SEL='pruned_equals_dense or pruned_non_finite or cluster_sizes_agree or ffps_adversarial or ffps_full_batch or dfps_cluster_stress or group_concat or three_interpolate'
timeout -k 10 1200 compute-sanitizer --tool synccheck --error-exitcode 77 --log-file $OUT/sanitizer_r3y_synccheck.log \
   python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 1000 -k "$SEL" -p no:cacheprovider > $OUT/sanitizer_r3y_synccheck_pytest.log 2>&1
echo "synccheck rc=$?"; tail -2 $OUT/sanitizer_r3y_synccheck_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY" $OUT/sanitizer_r3y_synccheck.log | tail -2
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_r3y.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_r3y.log | cut -c1-300
echo "== bench"; timeout -k 10 600 python bench.py --no-extras --no-cpu-baseline > $OUT/bench_r3y.json 2> $OUT/bench_r3y.err; echo "bench rc=$?"; head -c 250 $OUT/bench_r3y.json; echo
