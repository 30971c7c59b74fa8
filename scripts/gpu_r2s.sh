#!/bin/bash
# Round 2 sanitizer call: memcheck + racecheck + synccheck on the kernels added / changed in round 2.
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
SEL='full_pose or shared_grid or group_concat or gather_xyz or wrapper_golden or small_cloud'
for tool in memcheck racecheck synccheck; do
  timeout -k 10 900 compute-sanitizer --tool $tool --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_$tool.log \
     python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 800 -k "$SEL" -p no:cacheprovider > $OUT/sanitizer_${TAG}_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 $OUT/sanitizer_${TAG}_${tool}_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_${TAG}_$tool.log | tail -2
done
echo "== sa_mlp under memcheck (tcgen05 / TMEM kernels; small shapes)"
timeout -k 10 600 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_samlp_memcheck.log \
   python -m pytest tests/test_sa_fused_gpu.py -m gpu -q -x --timeout 500 -k "8x11 or 16x35 or 32x4 or 16x131" -p no:cacheprovider > $OUT/sanitizer_${TAG}_samlp_pytest.log 2>&1
echo "sa_mlp memcheck rc=$?"; tail -2 $OUT/sanitizer_${TAG}_samlp_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY" $OUT/sanitizer_${TAG}_samlp_memcheck.log | tail -2
echo "== dropin tests (incl. backbone)"; timeout -k 10 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q --timeout 300 > $OUT/pytest_dropin_${TAG}.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_dropin_${TAG}.log | cut -c1-200
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_${TAG}.log | cut -c1-300
