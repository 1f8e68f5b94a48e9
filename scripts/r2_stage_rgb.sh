mkdir -p gpurun_out
STAGE_MODES="rgb" bash scripts/stage_capture.sh 2>&1 | tail -4
