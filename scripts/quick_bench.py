"""Exploration helper (not the bench contract): evals/s of the host-buffer batch call at several chunk sizes."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ncand = 64
chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32, 64, 128, 256]
for name, kw in [("rgb", {}), ("dither", {"dither": True}), ("lab", {"perceptual_palettes": True, "subpalette_count": 4, "subpalette_size": 7})]:
    cfg = engine.Config(**{"subpalette_count": 8, "subpalette_size": 15, **kw})
    ctx = engine.Context(0)
    t = time.time()
    imgs = [engine.OptimizedImage(ctx, synth.image(s, "V"), cfg) for s in range(nimg)]
    t1 = time.time()
    engine.batch_initialize_tiles(imgs)
    t2 = time.time()
    engine.batch_recalculate_palettes(imgs)
    t3 = time.time()
    print(f"[{name}] create {t1 - t:.3f}s init_tiles {t2 - t1:.3f}s recalc {t3 - t2:.3f}s", flush=True)
    cand = np.stack([synth.candidates(s, 0, ncand) for s in range(nimg)])
    for ch in chunks:
        ctx.set_chunk(ch)
        engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
        t = time.time()
        reps = 3
        for _ in range(reps):
            engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
        dt = (time.time() - t) / reps
        print(f"[{name}] chunk {ch:4d}: {dt * 1e3:8.2f} ms/step  {nimg * ncand / dt:10.0f} evals/s", flush=True)
    for im in imgs:
        im.close()
    ctx.close()
