# one gpurun call: ncu --set full of the kernels around the scorer, per mode (the scorer itself is in final_capture.sh);
# 16 images x 64 candidates so that the image-creation launches do not use up the launch count
K='regex:k_assign|k_pyramid<0>|k_pool_fused|k_argmin|k_kmeans|k_tile_means|k_centres_to_palette'
for mode in rgb lab dither; do
  timeout 100 python scripts/quick_bench.py 16 $mode v3 > gpurun_out/s_$mode.log 2>&1 || exit 1
  timeout 200 ncu --set full --clock-control none --kernel-name-base demangled -k "$K" -c 90 -f -o gpurun_out/s_$mode python scripts/quick_bench.py 16 $mode v3 > gpurun_out/s_ncu_$mode.log 2>&1
  echo "$mode rc=$?"
done
ls -la gpurun_out/s_*.ncu-rep
