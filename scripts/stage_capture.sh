# ncu --set full of the kernels around the scorer, per mode (the scorer itself is captured by scripts/r2_capture.sh).
# 4 images x 64 candidates per launch keeps the per-evaluation buffers ncu saves and restores around every replayed pass small.
# Only the raw-page CSV of each report travels back (the .ncu-rep files are 35-40 MB each; gpurun_out is capped at 64 MiB).
K='regex:k_assign_pyr|k_assign_prepare|k_assign_dither|k_assign_rgb|k_assign_lab|k_pyramid|k_kmeans|k_tile_means|k_gather_points|k_pool_fused|k_argmin|k_tables'
for mode in ${STAGE_MODES:-rgb lab dither}; do
  timeout 100 python scripts/quick_bench.py 4 $mode v3 > gpurun_out/s_$mode.log 2>&1 || { echo "$mode plain run failed"; tail -5 gpurun_out/s_$mode.log; continue; }
  timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k "$K" -c 16 -f -o /tmp/s_$mode python scripts/quick_bench.py 4 $mode v3 > gpurun_out/s_ncu_$mode.log 2>&1
  echo "$mode rc=$?"
  ncu -i /tmp/s_$mode.ncu-rep --page raw --csv > gpurun_out/s_${mode}_raw.csv 2>/dev/null
done
ls -la gpurun_out/s_*
# the dither kernel once more at 16 images x 64 candidates (1024 CTAs: ~7 of its 8 CTAs per SM resident), its working occupancy
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:k_assign_dither -s 1 -c 4 -f -o /tmp/s_dither16 python scripts/quick_bench.py 16 dither v3 > gpurun_out/s_ncu_dither16.log 2>&1
echo "dither16 rc=$?"
ncu -i /tmp/s_dither16.ncu-rep --page raw --csv > gpurun_out/s_dither16_raw.csv 2>/dev/null
