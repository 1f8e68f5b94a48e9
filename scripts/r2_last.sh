# cube-root guard moved to 2^120: self-check + plane parity, then the scorer's ncu capture again (roofline_traffic.json's hash covers common.cuh)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 200 python -m pytest tests -m gpu -x -q -k "cbrt or planes or zero_weight" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2l_tests.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r2l_plain.json 2> gpurun_out/r2l_plain.err; echo "plain rc=$?"; cut -c1-200 gpurun_out/r2l_plain.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_score_ -s 16 -c 2 -f -o gpurun_out/r2l_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r2l_ncu2.log 2>&1; echo "ncu2 rc=$?"
