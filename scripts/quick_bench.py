"""Exploration helper (not the bench contract): device-timed evals/s of the batch evaluation for scorer variants."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["rgb"]
ncand = 64
ALL = {"rgb": {}, "dither": {"dither": True}, "lab": {"perceptual_palettes": True, "subpalette_count": 4, "subpalette_size": 7},
       "nes": {"nes": True, "dither": True, "subpalette_count": 4, "subpalette_size": 3}}
for name in modes:
    cfg = engine.Config(**{"subpalette_count": 8, "subpalette_size": 15, **ALL[name]})
    ctx = engine.Context(0)
    imgs = [engine.OptimizedImage(ctx, synth.image(s, "V"), cfg) for s in range(nimg)]
    engine.batch_initialize_tiles(imgs)
    engine.batch_recalculate_palettes(imgs)
    cand = np.stack([synth.candidates(s, 0, ncand) for s in range(nimg)])
    variants = [("v3", 3, 32, 4096), ("v2", 2, 32, 4096)]
    if len(sys.argv) > 3:
        variants = [v for v in variants if v[0] in sys.argv[3].split(",")]
    for label, fused, bw, chunk in variants:
        ctx.set_scorer(fused, bw)
        ctx.set_chunk(chunk)
        engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
        reps = 3
        if "noprof" in sys.argv[4:]:   # wall time only: with every kernel timed on its own the library does not overlap launches
            reps = 10
            t = time.time()
            for _ in range(reps):
                engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
            dt = (time.time() - t) / reps
            print(f"[{name}] {label:13s}: wall {dt * 1e3:7.3f} ms/step unprofiled  {nimg * ncand / dt:9.0f} evals/s", flush=True)
            continue
        ctx.profile_begin()
        t = time.time()
        for _ in range(reps):
            engine.batch_eval_candidates(imgs, 0, 0, cand, want_scores=False)
        dt = (time.time() - t) / reps
        prof = ctx.profile_end()
        tot = sum(v["ms"] for v in prof.values()) / reps
        top = ", ".join(f"{k}={v['ms'] / reps:.2f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:4])
        print(f"[{name}] {label:13s}: wall {dt * 1e3:7.2f} ms/step  kernels {tot:7.2f} ms  {nimg * ncand / (tot * 1e-3):9.0f} evals/s  | {top}", flush=True)
    for im in imgs:
        im.close()
    ctx.close()
