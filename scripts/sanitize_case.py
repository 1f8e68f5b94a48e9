"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel of the candidate path once, in all
four modes, on one image with a handful of candidates."""
import sys

import numpy as np

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

ctx = engine.Context(0)
for name, kw in {"rgb": {}, "lab": {"perceptual_palettes": True}, "dither": {"dither": True}, "nes": {"nes": True, "dither": True}}.items():
    cfg = engine.Config(subpalette_count=3, subpalette_size=4, **kw)
    im = engine.OptimizedImage(ctx, synth.image(5, "T"), cfg)
    im.initialize_tiles()
    im.recalculate_palettes()
    cand = synth.candidates(5, 0, 3)
    r = engine.batch_eval_candidates([im], 1, 2, cand[None])
    engine.batch_step_random([im], 1, 2, cand[None])
    # round 2's paths: several entries per launch, speculative iterations (fused finish, adopted maps), tile moves
    steps = [(0, 1), (2, 3), (1, 0)]
    m = engine.batch_eval_candidates_multi([im], steps, np.stack([synth.candidates(5, 1 + k, 3) for k in range(3)])[None])
    if kw.get("nes"):
        used = im.iterate("nes", steps[:2])
    else:
        used = im.iterate("random", steps, np.stack([synth.candidates(5, 9 + k, 3) for k in range(3)]))
        im.iterate("channel", [(1, 2, 0), (1, 2, 1)])
    engine.batch_step_tile_moves([im], np.asarray([[(3, 1), (40, 2)]], np.int32))
    print(name, r["scores"][0], m["best"]["idx"][0], used, im.error(), flush=True)
    im.close()
ctx.close()
