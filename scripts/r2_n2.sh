# two GPUs of one box: the two-GPU tests and the bench line as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2y_tests_n2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2y_tests_n2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2y_bench_n2.json; tail -2 gpurun_out/r2y_bench_n2.err
