#!/bin/bash
# Round 2, call E: full-pose IoU / NMS tests (+ whole suite), ops timing for the 9-DoF kernels.
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu_${TAG}.log
echo "== sanitizer 9dof"; timeout -k 10 600 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $OUT/sanitizer_${TAG}_memcheck.log python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "full_pose or shared_grid or group_concat or gather_xyz" -p no:cacheprovider > $OUT/sanitizer_${TAG}_pytest.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitizer_${TAG}_pytest.log; grep -E "ERROR SUMMARY" $OUT/sanitizer_${TAG}_memcheck.log | tail -2
