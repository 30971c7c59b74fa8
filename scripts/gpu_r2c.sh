#!/bin/bash
# Round 2, call C: group-kernel variants A/B, parity suite, new bench line (reference_cuda, other_configs), launch list.
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
echo "== group variants"; timeout -k 10 600 python scripts/group_variants.py run > $OUT/group_variants_${TAG}.txt 2>&1; echo "rc=$?"; cat $OUT/group_variants_${TAG}.txt | cut -c1-400
echo "== pytest gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_${TAG}.log
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_${TAG}.log
echo "== bench"; timeout -k 10 1200 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -5 $OUT/bench_${TAG}.err; head -c 300 $OUT/bench_${TAG}.json; echo
echo "== bench reference arm"; timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; echo "rc=$?"; head -c 300 $OUT/bench_ref_${TAG}.json; echo
