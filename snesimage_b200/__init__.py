"""snesimage_b200 -- B200-native engine for the palette-optimisation hot path of aexoden/snesimage.

csrc/       hand-written sm_100a CUDA kernels + the C ABI (include/snesgpu.h) -> libsnesgpu.so
engine.py   host-side mirror of the reference's OptimizedImage over that ABI (ctypes)
driver.py   headless optimiser schedule and the candidate-sharded multi-GPU step
synth.py    seeded synthetic images and candidate lists
"""
from .engine import Config, Context, OptimizedImage  # noqa: F401
