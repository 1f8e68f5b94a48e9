"""Synthetic inputs for the palette-optimisation hot path.

Counter-based, integer-only generators built on the splitmix64 finaliser, so the same bytes can be
reproduced in any language (tests/golden/gen_reference_vectors.rs restates `mix64`/`hashn` in Rust).  The reference
itself ships no sample image and draws its candidates from an unseeded `rand::rng()`
(/root/reference/src/lib.rs:201-208); here the candidate list is an explicit, seeded input.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)


def mix64(x):
    """splitmix64 output function on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=np.uint64) + _GOLD
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        return z ^ (z >> np.uint64(31))


def hashn(seed, *vals):
    """h = mix(seed); for v in vals: h = mix(h ^ (v * GOLD)).  Broadcasts over array arguments."""
    with np.errstate(over="ignore"):
        h = mix64(np.uint64(seed))
        for v in vals:
            h = mix64(h ^ (np.asarray(v).astype(np.uint64) * _GOLD))
        return h


def _value_noise(seed, channel):
    """4-octave value noise in 0..255, integer bilinear interpolation, octave weights 8/4/2/1."""
    y, x = np.mgrid[0:256, 0:256].astype(np.int64)
    acc = np.zeros((256, 256), np.int64)
    for o, (cell, wgt) in enumerate(((64, 8), (32, 4), (16, 2), (8, 1))):
        gx, gy, fx, fy = x // cell, y // cell, x % cell, y % cell

        def lat(ix, iy):
            return (hashn(seed, 1000 + channel, o, ix, iy) & np.uint64(0xFF)).astype(np.int64)

        v = (lat(gx, gy) * (cell - fx) * (cell - fy) + lat(gx + 1, gy) * fx * (cell - fy) +
             lat(gx, gy + 1) * (cell - fx) * fy + lat(gx + 1, gy + 1) * fx * fy) // (cell * cell)
        acc += wgt * v
    # the octave average hugs mid-grey; stretch x2.5 about 128 so the image spans the gamut
    return 128 + ((acc // 15) - 128) * 5 // 2


def _noise(seed, channel, amp):
    y, x = np.mgrid[0:256, 0:256].astype(np.int64)
    return (hashn(seed, 2000 + channel, x, y) % np.uint64(2 * amp + 1)).astype(np.int64) - amp


def image(seed: int, family: str = "V") -> np.ndarray:
    """256x256 RGBA8 synthetic image.  Families: G gradients, V value noise, B flat tile regions,
    T = V with ~5% fully transparent 8x8 tiles (alpha 0, RGB kept)."""
    y, x = np.mgrid[0:256, 0:256].astype(np.int64)
    out = np.zeros((256, 256, 4), np.int64)
    out[..., 3] = 255
    if family in ("V", "T"):
        for c in range(3):
            out[..., c] = _value_noise(seed, c) + _noise(seed, c, 2)
    elif family == "G":
        base = (x, y, (x + y) // 2)
        for c in range(3):
            out[..., c] = base[c] + _noise(seed, c, 8)
    elif family == "B":
        region = (hashn(seed, 3000, x // 32, y // 32) % np.uint64(8)).astype(np.int64)
        for c in range(3):
            col = (hashn(seed, 3001 + c, region) % np.uint64(240)).astype(np.int64) + 8
            out[..., c] = col + _noise(seed, c, 2)
    else:
        raise ValueError(f"unknown image family {family!r}")
    if family == "T":
        transparent = (hashn(seed, 4000, x // 8, y // 8) % np.uint64(100)) < np.uint64(5)
        out[..., 3] = np.where(transparent, 0, 255)
    return np.clip(out, 0, 255).astype(np.uint8)


def candidates(seed: int, iteration: int, n: int = 64) -> np.ndarray:
    """The n uniform BGR555 trial colours of one `optimize_palette_entry_random` call
    (lib.rs:205-208): (r, g, b) each uniform in 0..32, as an (n, 3) uint8 array."""
    k = np.arange(n, dtype=np.uint64)
    h = hashn(seed, 5000, iteration, k)
    return np.stack([(h & np.uint64(31)), ((h >> np.uint64(5)) & np.uint64(31)),
                     ((h >> np.uint64(10)) & np.uint64(31))], axis=1).astype(np.uint8)


def random_palette(seed: int, sub_count: int, sub_size: int) -> np.ndarray:
    """Uniform random 5-bit palette, (sub_count*sub_size, 3) uint8 — for kernel-level tests."""
    k = np.arange(sub_count * sub_size, dtype=np.uint64)
    h = hashn(seed, 6000, k)
    return np.stack([(h & np.uint64(31)), ((h >> np.uint64(5)) & np.uint64(31)),
                     ((h >> np.uint64(10)) & np.uint64(31))], axis=1).astype(np.uint8)


def random_tile_palettes(seed: int, sub_count: int) -> np.ndarray:
    k = np.arange(1024, dtype=np.uint64)
    return (hashn(seed, 7000, k) % np.uint64(sub_count)).astype(np.uint8)
