#!/bin/bash
# compute-sanitizer over a representative slice of the GPU parity suite (memcheck + racecheck + synccheck).
mkdir -p gpurun_out
SEL='golden or radius_boundary or fused_ffps_equals or fused_ffps_adversarial or group_concat or nms_batched or class_agnostic or dist_matrix_vs or sfps or chain_vs_oracle or cluster_large or ball_query_large or multi_sample or points_in_boxes3d'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --log-file gpurun_out/sanitizer_$tool.log \
     python -m pytest tests -m gpu -q -x --timeout 1400 -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|errors" gpurun_out/sanitizer_$tool.log | tail -3
done
