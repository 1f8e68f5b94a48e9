"""BASELINE.json configs[0..3] end to end on one GPU: synthetic 256x256 image (family V, seed 0), k-means init + tile
assignment + 100 iterations of the schedule of run() through the headless driver (one image: the way the reference
itself is used).  Prints wall-clock time, candidate evaluations per second and the error trajectory endpoints.

  cfg1  8 x 15, RGB distance, no dither
  cfg2  --perceptual-palettes, 4 x 7
  cfg3  --dither, 8 x 15
  cfg4  --nes --dither, 4 x 3
"""
import sys
import time

sys.path.insert(0, ".")
from snesimage_b200 import driver, engine, synth

CFGS = {
    "cfg1": dict(subpalette_count=8, subpalette_size=15),
    "cfg2": dict(subpalette_count=4, subpalette_size=7, perceptual_palettes=True),
    "cfg3": dict(subpalette_count=8, subpalette_size=15, dither=True),
    "cfg4": dict(subpalette_count=4, subpalette_size=3, nes=True, dither=True),
}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = engine.Context(0)
rgba = synth.image(0, "V")
for name, kw in CFGS.items():
    cfg = engine.Config(**kw)
    r = driver.HeadlessRunner(ctx, rgba, cfg, seed=0, ncand=64)
    t0 = time.perf_counter()
    r.initialize()
    ctx.synchronize()
    t1 = time.perf_counter()
    e0 = r.image.error()
    t2 = time.perf_counter()
    r.iterate(iters)
    ctx.synchronize()
    t3 = time.perf_counter()
    per_iter = 56 if cfg.nes else 64      # candidates per iteration (all iterations are `random` for < 4 sweeps, lib.rs:890)
    evals = iters * (per_iter + 2)        # + the bookkeeping optimize()/error() pairs of lib.rs:199/237 and 906-910
    print(f"{name}: init {1e3 * (t1 - t0):7.1f} ms | {iters} iterations {1e3 * (t3 - t2):8.1f} ms = {1e3 * (t3 - t2) / iters:6.2f} ms/iteration, "
          f"{evals / (t3 - t2):9.0f} evals/s | error {e0:.6f} -> {r.image.error():.6f}", flush=True)
    r.image.close()
ctx.close()
