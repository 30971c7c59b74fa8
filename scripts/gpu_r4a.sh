#!/bin/bash
# call 4a: after mbarrier.inval at kernel exit -- synccheck on the cluster / TMA kernels incl. multi-wave launches, whole suite, smoke, bench.
OUT=gpurun_out
SEL='pruned_equals_dense or pruned_non_finite or cluster_sizes_agree or ffps_adversarial or ffps_full_batch or dfps_cluster_stress or group_concat or three_interpolate or group_points'
timeout -k 10 1200 compute-sanitizer --tool synccheck --error-exitcode 77 --log-file $OUT/sanitizer_r4a_synccheck.log \
   python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 1000 -k "$SEL" -p no:cacheprovider > $OUT/sanitizer_r4a_synccheck_pytest.log 2>&1
echo "synccheck rc=$?"; tail -2 $OUT/sanitizer_r4a_synccheck_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY" $OUT/sanitizer_r4a_synccheck.log | tail -2
timeout -k 10 600 compute-sanitizer --tool synccheck --error-exitcode 77 --log-file $OUT/sanitizer_r4a_samlp_synccheck.log \
   python -m pytest tests/test_sa_fused_gpu.py -m gpu -q -x --timeout 500 -k "8x11 or 16x35 or 131-128-256-256 or 32x4" -p no:cacheprovider > $OUT/sanitizer_r4a_samlp_pytest.log 2>&1
echo "sa_mlp synccheck rc=$?"; tail -2 $OUT/sanitizer_r4a_samlp_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY" $OUT/sanitizer_r4a_samlp_synccheck.log | tail -1
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_r4a.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_r4a.log | cut -c1-300
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_r4a.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_r4a.log | cut -c1-200
echo "== bench"; timeout -k 10 600 python bench.py --no-extras --no-cpu-baseline > $OUT/bench_r4a.json 2> $OUT/bench_r4a.err; echo "bench rc=$?"; head -c 250 $OUT/bench_r4a.json; echo
