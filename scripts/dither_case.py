"""One image, dithered optimize() a few times (single-CTA wavefront latency) and a 64-candidate evaluation."""
import sys
import time

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

ctx = engine.Context(0)
cfg = engine.Config(subpalette_count=8, subpalette_size=15, dither=True)
im = engine.OptimizedImage(ctx, synth.image(0, "V"), cfg)
im.initialize_tiles()
im.recalculate_palettes()
for _ in range(3):
    im.optimize()
ctx.synchronize()
t = time.perf_counter()
for _ in range(10):
    im.optimize()
ctx.synchronize()
print(f"optimize() with dither, one image: {(time.perf_counter() - t) * 100:.3f} ms per call")
ctx.profile_begin()
import numpy as np
engine.batch_eval_candidates([im], 0, 0, synth.candidates(0, 0, 64)[None])
print(ctx.profile_end())
