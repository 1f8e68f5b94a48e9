#!/bin/bash
# A/B of library variants: scripts/ab_n.sh <nimg> <mode> a b c ... (libsnesgpu_<x>.so), two interleaved rounds; prints wall ms per
# step (host clock around synchronous calls), the sum of the kernels' CUDA-event times (concurrent kernels count twice) and the top kernels
n=$1; mode=$2; shift; shift
for round in 1 2; do for v in "$@"; do echo -n "[$n img] $v: "; SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so python scripts/quick_bench.py $n $mode v3 2>&1 | tail -1 | sed 's/.*: wall *\([0-9.]* ms\)\/step *kernels *\([0-9.]* ms\).*|\(.*\)/wall \1 | kernels \2 |\3/'; done; done
