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
