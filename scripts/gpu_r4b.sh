#!/bin/bash
# call 4b: bench line (all legs) + CPU reference arm + ncu launch list of the final library of the round.
OUT=gpurun_out
echo "== bench"; timeout -k 10 900 python bench.py > $OUT/bench_r4b.json 2> $OUT/bench_r4b.err; echo "bench rc=$?"; head -c 300 $OUT/bench_r4b.json; echo
echo "== bench reference arm"; timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_r4b.json 2> $OUT/bench_ref_r4b.err; echo "rc=$?"; head -c 200 $OUT/bench_ref_r4b.json; echo
echo "== ncu launches"; timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r4b.csv python scripts/ncu_chain.py > $OUT/ncu_launches_r4b.log 2>&1; echo "ncu rc=$?"
