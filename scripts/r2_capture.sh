# One gpurun call (1 GPU): GPU tests, smoke, bench, launch list, ncu --set full of the scorer, stage captures.
# usage: bash scripts/r2_capture.sh TAG [notests]
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*_launches.csv
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
fi
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_score_ -s 16 -c 2 -f -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu2 rc=$?"
[ "$3" = "nostages" ] || bash scripts/stage_capture.sh
du -sh gpurun_out
