# the scorer's ncu --set full capture again at HEAD (profiles/roofline_traffic.json is tied to a hash of the scorer's sources,
# common.cuh among them), then the bench line that reports it
TAG=${1:-r2s}
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err; echo "plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_score_ -s 16 -c 2 -f -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu2 rc=$?"
du -sh gpurun_out
