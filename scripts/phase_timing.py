"""Debug helper: per-phase cycle counts of k_score_v2 from an instrumented build (-DSNES_V2_TIMING).

build:  nvcc <flags of snesimage_b200/_build.py> -DSNES_V2_TIMING -o snesimage_b200/libsnesgpu_timing.so snesimage_b200/csrc/snesgpu.cu
run:    SNESGPU_SO=snesimage_b200/libsnesgpu_timing.so python scripts/phase_timing.py
"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

nimg, ncand = 16, 64
cfg = engine.Config(subpalette_count=8, subpalette_size=15)
ctx = engine.Context(0)
imgs = [engine.OptimizedImage(ctx, synth.image(s, "V"), cfg) for s in range(nimg)]
engine.batch_initialize_tiles(imgs)
engine.batch_recalculate_palettes(imgs)
cand = np.stack([synth.candidates(s, 0, ncand) for s in range(nimg)])
L = engine.lib()
buf = (C.c_ulonglong * 16)()
engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
L.snes_debug_v2_timing(buf, 1)
engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
L.snes_debug_v2_timing(buf, 1)
v = np.array(list(buf), dtype=np.float64)
ctas = nimg * ncand * 3
names = ["stage", "H", "V", "maps"] if engine.os.environ.get("SNESGPU_FUSED", "3") == "3" else ["stage", "H", "V warp busy", "maps warp busy", "V+maps phase"]
for base, label in ((0, "scale 0"), (8, "scales 1-5")):
    tot = v[base] + v[base + 1] + (v[base + 2] + v[base + 3] if len(names) == 4 else v[base + 4])
    print(f"{label}: cycles per CTA {tot / ctas:10.0f}")
    for i, n in enumerate(names):
        print(f"   {n:16s} {v[base + i] / ctas:10.0f}  ({100 * v[base + i] / tot:5.1f}%)")
