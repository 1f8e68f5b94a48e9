"""Shared helpers of the parity tests: build the same OptimizedImage state in the oracle and on the GPU."""
import numpy as np

from oracle import binding as ob
from snesimage_b200 import engine, synth


def make_pair(ctx, rgba, C, S, dither=False, lab=False, nes=False, seed=1, random_state=True):
    cfg = engine.Config(subpalette_count=C, subpalette_size=S, dither=dither, perceptual_palettes=lab, nes=nes)
    g = engine.OptimizedImage(ctx, rgba, cfg)
    o = ob.OracleImage(rgba, C, S, dither, lab, nes)
    if random_state:
        pal = synth.random_palette(seed, C, S)
        if nes:
            pal = np.stack([ob.nes_color(int(v) % 56) for v in synth.hashn(seed, 9000, np.arange(C * S))])
        tp = synth.random_tile_palettes(seed, C)
        for im in (g, o):
            im.palette = pal
            im.tile_palettes = tp
    return g, o


def lab_choice_ok(o, gmap, tol=2e-4):
    """Every GPU choice must be (within tol) as close as the oracle's own choice, measured with the
    oracle's CIEDE2000.  Returns (number of differing pixels, worst excess distance)."""
    omap = o.palette_map
    diff = np.argwhere(omap != gmap)
    worst = 0.0
    pal = o.palette
    tp = o.tile_palettes
    for y, x in diff:
        if o.rgba[y, x, 3] == 0:
            return len(diff), float("inf")  # transparent pixels must both be 0
        sub = int(tp[(y // 8) * 32 + x // 8]) * o.sub_size
        t = o.rgba[y, x, :3]
        dg = ob.cielab(ob.snes_as_rgba(pal[sub + gmap[y, x]])[:3], t)
        do = ob.cielab(ob.snes_as_rgba(pal[sub + omap[y, x]])[:3], t)
        worst = max(worst, dg - do)
    return len(diff), worst


# ---- the oracle's candidate loop, farmed over the host cores ---------------------------------------------------------
_W = {}


def _pool_init(rgba, cfg_tuple):
    _W["o"] = ob.OracleImage(rgba, *cfg_tuple)


def _pool_eval(task):
    palette, tile_palettes, p, i, cand = task
    o = _W["o"]
    o.tile_palettes = tile_palettes
    o.palette = palette
    return o.eval_candidates(p, i, cand)


class OraclePool:
    """One oracle image per worker process; `eval` scores a candidate list (entry (p, i) := cand[k]; optimize(); error(),
    lib.rs:209-214) against the given state, split over the workers.  Spawned, not forked: the test process holds a CUDA
    context."""

    def __init__(self, rgba, cfg, procs=None):
        import multiprocessing as mp
        import os
        self.procs = procs or max(1, min(32, len(os.sched_getaffinity(0))))
        tup = (cfg.subpalette_count, cfg.subpalette_size, bool(cfg.dither), bool(cfg.perceptual_palettes), bool(cfg.nes))
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_pool_init, initargs=(np.ascontiguousarray(rgba), tup))

    def eval(self, palette, tile_palettes, p, i, cand) -> np.ndarray:
        cand = np.ascontiguousarray(cand, np.uint8).reshape(-1, 3)
        parts = [c for c in np.array_split(cand, min(self.procs, len(cand))) if len(c)]
        res = self.pool.map(_pool_eval, [(palette, tile_palettes, p, i, c) for c in parts], chunksize=1)
        return np.concatenate(res)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.pool.terminate()
        self.pool.join()


def oracle_entry_step(o, pool, mode, p, i, channel, cand):
    """optimize_palette_entry_{random,channel,nes} (lib.rs:191-240, 286-328, 242-284) rebuilt on eval_candidates so that
    the candidate loop can run on a pool: the reference's sequential strict-< scan ends on the FIRST minimum of the
    candidates' errors, and takes it only if it beats the starting error (f64::MAX in NES mode)."""
    S = o.sub_size
    slot = p * S + i
    pal = o.palette
    if mode == "nes":
        cands = np.stack([ob.nes_color(k) for k in range(56)])          # lib.rs:252-253
        best_error = np.finfo(np.float64).max                           # lib.rs:250
    else:
        best_error = o.error()                                          # lib.rs:199, 294
        if mode == "random":
            cands = np.ascontiguousarray(cand, np.uint8).reshape(-1, 3)
        else:
            cands = np.repeat(pal[slot][None], 32, axis=0)              # lib.rs:296-297
            cands[:, channel] = np.arange(32)
    scores = pool.eval(pal, o.tile_palettes, p, i, cands) if pool is not None else o.eval_candidates(p, i, cands)
    k = int(np.argmin(scores))                                          # numpy argmin = first minimum
    if scores[k] < best_error:
        pal[slot] = cands[k]
    o.palette = pal
    o.optimize()                                                        # lib.rs:236-237, 280-281, 324-325
    return scores
