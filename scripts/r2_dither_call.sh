# one gpurun call: dither parity tests with the library in the tree, then an A/B of dither kernel variants (dither + nes modes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dither or nes or tile_move or multi_entry or invalid or iterate" > gpurun_out/rd_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/rd_tests.log
for v in "$@"; do for m in dither nes; do echo -n "$v: "; SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 120 python scripts/quick_bench.py 64 $m v3 2>&1 | tail -1; done; done | tee gpurun_out/rd_ab.log
timeout 900 python -m pytest tests/test_trajectory.py -m gpu -x -q -k "cfg3 or cfg4" > gpurun_out/rd_traj.log 2>&1; echo "traj rc=$?"; tail -4 gpurun_out/rd_traj.log
