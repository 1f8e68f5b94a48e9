#!/bin/bash
# A/B of library variants on one box for one mode: scripts/ab_mode.sh <mode> a b c ... (libsnesgpu_<x>.so), two rounds, interleaved
mode=$1; shift
for round in 1 2; do for v in "$@"; do echo -n "$v: "; SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so python scripts/quick_bench.py 64 $mode v3 2>&1 | tail -1 | sed 's/.*kernels *\([0-9.]* ms\).*|\(.*\)/\1 |\2/'; done; done
