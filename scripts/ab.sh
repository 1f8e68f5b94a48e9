#!/bin/bash
# A/B of library variants on one box: scripts/ab.sh a b c ... (libsnesgpu_<x>.so), two rounds each, interleaved
for round in 1 2; do for v in "$@"; do echo -n "$v: "; SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so python scripts/quick_bench.py 64 rgb v3 2>&1 | tail -1 | sed 's/.*k_score_v3=\([0-9.]*\).*/\1/'; done; done
