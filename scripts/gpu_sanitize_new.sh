#!/bin/bash
# compute-sanitizer (memcheck + racecheck + synccheck) over the tests of the kernels changed late in round 1:
# 6-/8-CTA cluster F-FPS, two-phase NMS / IoU matrix, input staging.
mkdir -p gpurun_out
SEL='small_cloud or cluster_sizes_agree or nms_vs_oracle or nms_batched or iou_matrices or class_agnostic or break_up_pc or stage_frames'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --log-file gpurun_out/sanitizer2_$tool.log \
     python -m pytest tests -m gpu -q -x --timeout 800 -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer2_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer2_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|errors" gpurun_out/sanitizer2_$tool.log | tail -3
done
