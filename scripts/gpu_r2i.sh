#!/bin/bash
# bench only (+ fused-SA ncu capture, small)
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
echo "== bench"; timeout -k 10 1200 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -5 $OUT/bench_${TAG}.err; head -c 300 $OUT/bench_${TAG}.json; echo
echo "== ncu full, fused SA"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:sa_mlp_kernel" -c 6 -f -o /tmp/prof_samlp_$TAG python scripts/ncu_sa_mlp.py > $OUT/prof_samlp_$TAG.log 2>&1; echo "ncu rc=$?"; ncu -i /tmp/prof_samlp_$TAG.ncu-rep --page raw --csv > $OUT/prof_samlp_${TAG}_raw.csv 2>/dev/null; ncu -i /tmp/prof_samlp_$TAG.ncu-rep --page source --csv > $OUT/prof_samlp_${TAG}_source.csv 2>/dev/null; ls -la $OUT/prof_samlp_${TAG}_*.csv; du -sh $OUT
