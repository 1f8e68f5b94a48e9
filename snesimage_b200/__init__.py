"""snesimage_b200 -- B200-native engine for the palette-optimisation hot path of aexoden/snesimage.

csrc/       hand-written sm_100a CUDA kernels + the C ABI (include/snesgpu.h) -> libsnesgpu.so
engine.py   host-side mirror of the reference's OptimizedImage over that ABI (ctypes)
driver.py   headless optimiser schedule and the candidate-sharded multi-GPU step
synth.py    seeded synthetic images and candidate lists
ingest.py   image decode (PIL) with the reference's size check; JSON -> optimiser state (resume)
__main__.py the reference's command line over the headless driver: python -m snesimage_b200 SRC DST -c 8 -s 15 ...
"""
from .engine import Config, Context, OptimizedImage  # noqa: F401
