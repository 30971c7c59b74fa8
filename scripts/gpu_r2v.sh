#!/bin/bash
# Round 2 call V: ncu source-level captures of the pruned / dense fused F-FPS kernel (one wave of 6-CTA clusters).
OUT=gpurun_out
mkdir -p $OUT
for cfg in "2 6 0.1 pruned_f01" "1 6 1.0 dense_f10"; do
  set -- $cfg
  DE6D_PRUNE=$1 DE6D_S=$2 DE6D_FSCALE=$3 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:fps_features" -s 1 -c 1 -f -o $OUT/prof_ffps_$4 python scripts/ncu_ffps2.py > $OUT/prof_ffps_$4.log 2>&1
  echo "ncu $4 rc=$?"; tail -2 $OUT/prof_ffps_$4.log | cut -c1-200; ls -la $OUT/prof_ffps_$4.ncu-rep
done
