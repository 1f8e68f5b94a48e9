#!/bin/bash
# builds the product library and (with "timing") the instrumented debug variant used by scripts/phase_timing.py
set -e
cd /root/repo/snesimage_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared"
nvcc $FLAGS -Xptxas -v -o ../libsnesgpu.so snesgpu.cu 2>&1 | grep -A2 "k_score_v2\|k_score_v3\|error" | grep -v "^--" | tail -6
if [ "$1" = "timing" ]; then shift; nvcc $FLAGS -DSNES_V2_TIMING "$@" -o ../libsnesgpu_timing.so snesgpu.cu 2>&1 | grep -i "error" || true; fi
