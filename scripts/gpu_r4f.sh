#!/bin/bash
# call 4f: memcheck + synccheck on the D-FPS kernels after the last changes of the round
OUT=gpurun_out
SEL='multi_sample_rounds or dfps_cluster_stress or fps_full_size or sfps or small_cloud'
for tool in memcheck synccheck; do
  timeout -k 10 600 compute-sanitizer --tool $tool --error-exitcode 77 --log-file $OUT/sanitizer_r4f_$tool.log \
     python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 500 -k "$SEL" -p no:cacheprovider > $OUT/sanitizer_r4f_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -1 $OUT/sanitizer_r4f_${tool}_pytest.log | cut -c1-200; grep -E "ERROR SUMMARY" $OUT/sanitizer_r4f_$tool.log | tail -1
done
