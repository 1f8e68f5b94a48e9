# one gpurun call: interleaved A/B of scorer variants, then the scorer-facing parity tests on every variant
bash scripts/ab.sh "$@" 2>&1 | tee gpurun_out/ab.log
for v in "$@"; do
  [ "$v" = base ] && continue
  echo "== parity $v" | tee -a gpurun_out/ab.log
  SNESGPU_SO=snesimage_b200/libsnesgpu_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "error or scorer or candidates or batch or random" 2>&1 | tail -2 | tee -a gpurun_out/ab.log
done
