"""CPU checks of the helpers the trajectory parity tests stand on (tests/util.py): the candidate loop rebuilt on the
oracle's eval_candidates, and its farming over worker processes."""
import numpy as np

from oracle import binding as ob
from snesimage_b200 import engine, synth
from util import OraclePool, oracle_entry_step


def _oracle(rgba, cfg):
    return ob.OracleImage(rgba, cfg.subpalette_count, cfg.subpalette_size, cfg.dither, cfg.perceptual_palettes, cfg.nes)


def test_python_entry_step_equals_the_oracles_own():
    """`oracle_entry_step` (the candidate loop of lib.rs:191-328 rebuilt on the oracle's eval_candidates so that it can be
    farmed out) gives what the oracle's own optimize_palette_entry_{random,channel,nes} give.  CPU only, small palettes."""
    rgba = synth.image(5, "B")
    for kw, mode in [(dict(subpalette_count=2, subpalette_size=3), "random"), (dict(subpalette_count=2, subpalette_size=3), "channel"),
                     (dict(subpalette_count=2, subpalette_size=2, nes=True, dither=True), "nes")]:
        cfg = engine.Config(**kw)
        a, b = _oracle(rgba, cfg), _oracle(rgba, cfg)
        for o in (a, b):
            o.initialize_tiles()
            o.recalculate_palettes()
        cand = synth.candidates(1, 0, 6)
        for p, i in [(0, 1), (1, 0)]:
            if mode == "random":
                a.optimize_palette_entry_random(p, i, cand)
            elif mode == "channel":
                a.optimize_palette_entry_channel(p, i, 1)
            else:
                a.optimize_palette_entry_nes(p, i)
            oracle_entry_step(b, None, mode, p, i, 1, cand)
            assert np.array_equal(a.palette, b.palette) and np.array_equal(a.palette_map, b.palette_map), (mode, p, i)



def test_oracle_pool_equals_direct_evaluation():
    rgba = synth.image(6, "T")
    cfg = engine.Config(subpalette_count=2, subpalette_size=3, dither=True)
    o = _oracle(rgba, cfg)
    o.initialize_tiles()
    o.recalculate_palettes()
    cand = synth.candidates(2, 0, 5)
    want = o.eval_candidates(1, 2, cand)
    with OraclePool(rgba, cfg, procs=2) as pool:
        got = pool.eval(o.palette, o.tile_palettes, 1, 2, cand)
        again = pool.eval(o.palette, o.tile_palettes, 1, 2, cand[:1])
    assert np.array_equal(got, want) and again[0] == want[0]
