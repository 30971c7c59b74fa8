#!/bin/bash
OUT=gpurun_out
for c in waves six four; do
  timeout -k 10 300 compute-sanitizer --tool synccheck --error-exitcode 77 --log-file $OUT/sync_case_$c.log python scripts/synccheck_ffps_cases.py $c > $OUT/sync_case_${c}_out.log 2>&1
  echo "$c rc=$?"; tail -1 $OUT/sync_case_${c}_out.log | cut -c1-150; grep -E "ERROR SUMMARY" $OUT/sync_case_$c.log | head -1; grep -m1 "by thread" $OUT/sync_case_$c.log
done
