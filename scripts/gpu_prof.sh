#!/bin/bash
# ncu --set full capture of named kernels from the bench chain (eager, single stream).  Usage: gpu_prof.sh <tag> <kernel regex> [count]
TAG=$1; PAT=$2; CNT=${3:-2}
mkdir -p gpurun_out
DE6D_STEPS=0 timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -c $CNT -f -o gpurun_out/prof_$TAG python scripts/ncu_chain.py > gpurun_out/prof_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_$TAG.log; ls -la gpurun_out/prof_$TAG.ncu-rep
