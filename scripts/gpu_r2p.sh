#!/bin/bash
# Round 2 profiling call: ncu launch list + full capture of one chain step (exports only; the .ncu-rep stays in /tmp)
TAG=${1:-r2p}
bash scripts/gpu_prof_r2.sh $TAG
