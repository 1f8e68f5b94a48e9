mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lab or cielab or multi_entry or fuzz or kmeans or initialize or tile_move" > gpurun_out/rd3_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/rd3_tests.log
timeout 120 python scripts/quick_bench.py 64 lab v3 2>&1 | tail -1 | tee gpurun_out/rd3_lab.log
timeout 200 python scripts/kmeans_time.py 2>&1 | tail -8 | tee gpurun_out/rd3_kmeans.log
