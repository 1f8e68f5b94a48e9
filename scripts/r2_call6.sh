mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dither or nes or split or trajectory_rgb" > gpurun_out/rd6_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/rd6_tests.log
for r in 1 2; do for v in nosplit HEAD; do for m in dither nes; do echo -n "$v: "; if [ $v = HEAD ]; then timeout 120 python scripts/quick_bench.py 64 $m v3 noprof 2>&1 | tail -1; else SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 120 python scripts/quick_bench.py 64 $m v3 noprof 2>&1 | tail -1; fi; done; done; done | tee gpurun_out/rd6_ab.log
