#!/bin/bash
# builds scripts/micro/libfps_stats.so: the FPS kernels with the DE6D_FPS_STATS counters (tuning aid, not the product library)
cd "$(dirname "$0")/../.."
nvcc -shared -o scripts/micro/libfps_stats.so de6d_b200/csrc/fps.cu de6d_b200/csrc/fps_small.cu de6d_b200/csrc/capi.cu -DDE6D_FPS_STATS \
  -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -lcudart
