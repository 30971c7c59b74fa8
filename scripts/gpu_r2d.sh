#!/bin/bash
# Round 2, call D: tcgen05 bring-up test, group kernel (big-row mode) timings, bench line with reference_cuda, reference arm.
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
echo "== umma micro"; timeout -k 10 120 scripts/micro/umma_tf32 > $OUT/umma_${TAG}.txt 2>&1; echo "rc=$?"; cat $OUT/umma_${TAG}.txt
echo "== group variants"; timeout -k 10 600 python scripts/group_variants.py run > $OUT/group_variants_${TAG}.txt 2>&1; echo "rc=$?"; cat $OUT/group_variants_${TAG}.txt | cut -c1-400
echo "== pytest gpu (group, chain)"; timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -k "group or chain or gather" > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_${TAG}.log
echo "== bench"; timeout -k 10 1200 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -5 $OUT/bench_${TAG}.err; head -c 300 $OUT/bench_${TAG}.json; echo
echo "== bench reference arm"; timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; echo "rc=$?"; head -c 300 $OUT/bench_ref_${TAG}.json; echo
