#!/bin/bash
# Round 2 profiling call: ncu launch list of one serialised chain step, full ncu capture of one step (+ the fused SA kernel),
# raw CSV exports for profiles/.
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
echo "== ncu launches"; timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python scripts/ncu_chain.py > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu rc=$?"
echo "== ncu full, one step"; DE6D_STEPS=0 DE6D_TRACE=$OUT/prof_step_$TAG.trace.json timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k "regex:fps_|group_|gather_xyz|bq_grid|nms_kernel|ball_query_kernel" -c 60 -f -o /tmp/prof_step_$TAG python scripts/ncu_chain.py > $OUT/prof_step_$TAG.log 2>&1; echo "ncu rc=$?"; ncu -i /tmp/prof_step_$TAG.ncu-rep --page raw --csv > $OUT/prof_step_${TAG}_raw.csv 2>/dev/null; ls -la $OUT/prof_step_${TAG}_raw.csv /tmp/prof_step_$TAG.ncu-rep
echo "== ncu full, fused SA"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:sa_mlp_kernel" -c 6 -f -o /tmp/prof_samlp_$TAG python scripts/ncu_sa_mlp.py > $OUT/prof_samlp_$TAG.log 2>&1; echo "ncu rc=$?"; ncu -i /tmp/prof_samlp_$TAG.ncu-rep --page raw --csv > $OUT/prof_samlp_${TAG}_raw.csv 2>/dev/null; ls -la $OUT/prof_samlp_${TAG}_raw.csv
ncu -i /tmp/prof_samlp_$TAG.ncu-rep --page source --csv > $OUT/prof_samlp_${TAG}_source.csv 2>/dev/null; ls -la $OUT/prof_samlp_${TAG}_source.csv
du -sh $OUT
