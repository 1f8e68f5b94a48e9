#!/bin/bash
# builds snesimage_b200/libsnesgpu_<name>.so with extra nvcc flags: scripts/variant.sh <name> [-DFOO=1 ...]   (A/B runs via SNESGPU_SO)
set -e
name=$1; shift
cd "$(dirname "$0")/../snesimage_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared "$@" -o ../libsnesgpu_$name.so snesgpu.cu 2>&1 | grep -iE " error" || true
ls -la ../libsnesgpu_$name.so
