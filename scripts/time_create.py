import sys, time
sys.path.insert(0, ".")
from snesimage_b200 import engine, synth
t=time.perf_counter(); imgs=[synth.image(s,"V") for s in range(16)]; print("synth per image ms", (time.perf_counter()-t)/16*1e3)
ctx=engine.Context(0)
cfg=engine.Config(subpalette_count=8, subpalette_size=15)
a=engine.OptimizedImage(ctx, imgs[0], cfg)
t=time.perf_counter(); objs=[engine.OptimizedImage(ctx, im, cfg) for im in imgs]; ctx.synchronize(); print("snes_image_new per image ms", (time.perf_counter()-t)/16*1e3)
t=time.perf_counter(); engine.batch_initialize_tiles(objs); ctx.synchronize(); print("batch init tiles ms", (time.perf_counter()-t)*1e3)
t=time.perf_counter(); engine.batch_recalculate_palettes(objs); ctx.synchronize(); print("batch recalc ms", (time.perf_counter()-t)*1e3)
