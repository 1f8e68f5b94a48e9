set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?"
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
timeout 120 python scripts/quick_bench.py 64 rgb v3 > gpurun_out/g_quick.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/g_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_score_v3 -s 8 -c 1 -f -o gpurun_out/g_prof python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/g_ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/g_tests.log; cat gpurun_out/g_smoke.log | tail -2; cat gpurun_out/g_bench.json | cut -c1-400
