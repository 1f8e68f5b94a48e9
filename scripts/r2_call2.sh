mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dither or nes or kmeans or initialize or recalc or lab" > gpurun_out/rd2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/rd2_tests.log
for v in "$@"; do for m in dither nes; do echo -n "$v: "; SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 120 python scripts/quick_bench.py 64 $m v3 2>&1 | tail -1; done; done | tee gpurun_out/rd2_ab.log
for v in d2c6 d3c8; do echo "== kmeans $v"; SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 200 python scripts/kmeans_time.py 2>&1 | tail -8; done | tee gpurun_out/rd2_kmeans.log
