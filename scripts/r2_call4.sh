mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rd4_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/rd4_tests.log
for v in exactcbrt HEAD; do for m in rgb dither lab nes; do echo -n "$v: "; if [ $v = HEAD ]; then timeout 120 python scripts/quick_bench.py 64 $m v3 2>&1 | tail -1; else SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 120 python scripts/quick_bench.py 64 $m v3 2>&1 | tail -1; fi; done; done | tee gpurun_out/rd4_ab.log
