"""Randomised GPU-vs-oracle parity sweep (slow: the oracle costs ~50 ms per evaluation).

For a number of random configurations (image family, subpalette count / size, dither, NES, seed): k-means init on both
sides, then a few candidate evaluations and one optimiser step of each kind; integer outputs must be identical, errors
within 1e-8.  usage: python scripts/fuzz_parity.py [cases] [seed]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import binding as ob
from snesimage_b200 import engine, synth

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = engine.Context(0)
rng = np.random.default_rng(seed0)
bad = 0
t0 = time.time()
for case in range(cases):
    family = "VGBT"[int(rng.integers(4))]
    C = int(rng.integers(1, 9))
    S = int(rng.integers(2, 16))
    dither = bool(rng.integers(2))
    nes = bool(rng.integers(4) == 0)
    seed = int(rng.integers(1 << 20))
    rgba = synth.image(seed, family)
    cfg = engine.Config(subpalette_count=C, subpalette_size=S, dither=dither, nes=nes)
    g = engine.OptimizedImage(ctx, rgba, cfg)
    o = ob.OracleImage(rgba, C, S, dither, False, nes)
    tag = f"case {case}: {family} C={C} S={S} dither={dither} nes={nes} seed={seed}"
    try:
        try:
            o.initialize_tiles()
            o.recalculate_palettes()
        except RuntimeError:
            try:
                g.initialize_tiles()
                g.recalculate_palettes()
                print("MISMATCH (oracle refused k-means, GPU did not)", tag)
                bad += 1
            except engine.KmeansAssertion:
                pass
            continue
        g.initialize_tiles()
        g.recalculate_palettes()
        ok = np.array_equal(g.tile_palettes, o.tile_palettes) and np.array_equal(g.palette, o.palette) and np.array_equal(g.palette_map, o.palette_map)
        p, i = int(rng.integers(C)), int(rng.integers(S))
        cand = synth.candidates(seed, case, 5)
        sg = engine.batch_eval_candidates([g], p, i, cand[None])["scores"][0]
        so = o.eval_candidates(p, i, cand)
        ok = ok and np.max(np.abs(sg - so)) <= 1e-8
        if nes:
            g.optimize_palette_entry_nes(p, i)
            o.optimize_palette_entry_nes(p, i)
        else:
            g.optimize_palette_entry_random(p, i, cand)
            o.optimize_palette_entry_random(p, i, cand)
            ch = int(rng.integers(3))
            g.optimize_palette_entry_channel(p, i, ch)
            o.optimize_palette_entry_channel(p, i, ch)
        ok = ok and np.array_equal(g.palette, o.palette) and np.array_equal(g.palette_map, o.palette_map) and abs(g.error() - o.error()) <= 1e-8
        ok = ok and g.as_json() == o.as_json()
        if not ok:
            bad += 1
            print("MISMATCH", tag, flush=True)
    finally:
        g.close()
print(f"{cases} cases, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
