# last call of the round at HEAD: GPU tests, smoke, the bench line, the launch list (no ncu --set full: the scorer sources are unchanged since r2x)
TAG=${1:-r2y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu1 rc=$?"
STAGE_MODES="dither" bash scripts/stage_capture.sh > gpurun_out/${TAG}_stage.log 2>&1; tail -3 gpurun_out/${TAG}_stage.log
