# quick GPU check: GPU tests (-x), then the bench line with extras, then an A/B of library variants per mode
# usage: bash scripts/r2_quick.sh TAG "<pytest -k expr or empty>" "<variants>" "<modes>"
TAG=${1:-q}; KEXPR=$2; VARIANTS=$3; MODES=${4:-dither}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then timeout 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_tests.log 2>&1; else timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; fi
echo "tests rc=$?"; tail -15 gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4))
    for k, v in d.get("modes", {}).items(): print("mode", k, round(v["value"]), v["kernel_ms_per_step"])
    for k, v in d.get("configs", {}).items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "config"})
except Exception as e: print("bench line unreadable:", e)
PY
if [ -n "$VARIANTS" ]; then for m in $MODES; do bash scripts/ab_mode.sh $m $VARIANTS 2>&1 | tee -a gpurun_out/${TAG}_ab.log; done; fi
