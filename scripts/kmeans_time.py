"""Exploration helper: per-kernel time of initialize_tiles + recalculate_palettes for 1 and 64 pictures (SNESGPU_SO picks the library)."""
import sys

sys.path.insert(0, ".")
from snesimage_b200 import engine, synth

for mode, extra in (("rgb 8x15", {}), ("lab 4x7", {"perceptual_palettes": True, "subpalette_count": 4, "subpalette_size": 7})):
    for nimg in (1, 64):
        cfg = engine.Config(**{"subpalette_count": 8, "subpalette_size": 15, **extra})
        ctx = engine.Context(0)
        imgs = [engine.OptimizedImage(ctx, synth.image(s, "V"), cfg) for s in range(nimg)]
        for rep in range(2):
            ctx.profile_begin()
            engine.batch_initialize_tiles(imgs)
            engine.batch_recalculate_palettes(imgs)
            prof = ctx.profile_end()
        top = ", ".join(f"{k}={v['ms']:.3f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:5])
        print(f"[{mode}] {nimg:2d} pictures: {sum(v['ms'] for v in prof.values()):8.3f} ms | {top}", flush=True)
        pal = [im.palette.copy() for im in imgs[:2]]
        print("   palette checksum", [int(p.astype('int64').sum()) for p in pal])
        for im in imgs:
            im.close()
        ctx.close()
