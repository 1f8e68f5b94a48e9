mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cbrt or planes or eval_candidates or zero_weight" > gpurun_out/rd7_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/rd7_tests.log
for r in 1 2; do for v in prev HEAD; do for m in dither lab; do echo -n "$v: "; if [ $v = HEAD ]; then timeout 120 python scripts/quick_bench.py 64 $m v3 2>&1 | tail -1; else SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 120 python scripts/quick_bench.py 64 $m v3 2>&1 | tail -1; fi; done; done; done | tee gpurun_out/rd7_ab.log
