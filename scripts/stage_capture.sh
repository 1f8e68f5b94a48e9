# one gpurun call: ncu --set full of the kernels around the scorer, per mode (the scorer itself is in final_capture.sh).
# NOT YET RUN TO COMPLETION: round 1's only attempt (-c 90 launches per mode, k-means kernels included) ran into the call's
# time limit -- ncu saves and restores the GBs of per-evaluation buffers around every replayed pass -- and used up the
# round's last GPU minutes.  This version captures 12 launches per mode of the candidate-loop kernels only, 4 images.
K='regex:k_assign_pyr|k_assign_prepare|k_assign_dither|k_pyramid<0>|k_pool_fused|k_argmin'
for mode in rgb lab dither; do
  timeout 100 python scripts/quick_bench.py 4 $mode v3 > gpurun_out/s_$mode.log 2>&1 || exit 1
  timeout 150 ncu --set full --clock-control none --kernel-name-base demangled -k "$K" -c 12 -f -o gpurun_out/s_$mode python scripts/quick_bench.py 4 $mode v3 > gpurun_out/s_ncu_$mode.log 2>&1
  echo "$mode rc=$?"
done
ls -la gpurun_out/s_*.ncu-rep
