"""Cost of one snes_image_iterate call on ONE picture as a function of the speculation depth K (cfg1: 8 x 15, RGB): the
candidates are the entry's current colour, so nothing is ever accepted and every call stands for K iterations.
Prints wall ms per call, ms per iteration and the CUDA-event time of each kernel per call."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

kw = dict(subpalette_count=8, subpalette_size=15)
if len(sys.argv) > 1 and sys.argv[1] == "dither":
    kw["dither"] = True
ctx = engine.Context(0)
cfg = engine.Config(**kw)
im = engine.OptimizedImage(ctx, synth.image(0, "V"), cfg)
im.initialize_tiles()
im.recalculate_palettes()
pal = im.palette
for K in (1, 2, 3, 4, 6, 8, 12, 16):
    steps = [(k % 8, k % 15) for k in range(K)]
    cand = np.stack([np.repeat(pal[p * 15 + i][None], 64, axis=0) for p, i in steps])
    for _ in range(3):
        im.iterate("random", steps, cand)
    reps = 30
    ctx.profile_begin()
    t = time.perf_counter()
    for _ in range(reps):
        used, _, _ = im.iterate("random", steps, cand)
        assert used == K
    dt = (time.perf_counter() - t) / reps
    prof = ctx.profile_end()
    gpu = sum(v["ms"] for v in prof.values()) / reps
    top = ", ".join(f"{k}={v['ms'] / reps * 1e3:.0f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8])
    print(f"K={K:2d}: {dt * 1e3:6.3f} ms/call (profiled), {dt * 1e3 / K:6.3f} ms/iteration, kernels {gpu:6.3f} ms | us: {top}", flush=True)
    t = time.perf_counter()
    for _ in range(reps):
        im.iterate("random", steps, cand)
    dt = (time.perf_counter() - t) / reps
    print(f"      {dt * 1e3:6.3f} ms/call unprofiled = {64 * K / dt:9.0f} candidate evals/s", flush=True)
