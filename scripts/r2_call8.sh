mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "kmeans or initialize or recalc or nes or trajectory_rgb or invalid" > gpurun_out/rd8_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/rd8_tests.log
for v in 1 0; do echo "== SNESGPU_NO_CLUSTER_KMEANS=$v"; SNESGPU_NO_CLUSTER_KMEANS=$v timeout 200 python scripts/kmeans_time.py 2>&1 | grep -A1 "rgb"; done | tee gpurun_out/rd8_kmeans.log
