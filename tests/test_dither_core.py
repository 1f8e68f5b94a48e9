"""CPU checks of the dither kernel's per-pixel step (snesimage_b200/csrc/dither_core.h), compiled for the host.

The device kernel (dither.cuh) runs dc::step from that header; tests/host/dither_emulate.cpp runs the same function for 128
emulated threads.  Here it is compared with the oracle's optimize() (lib.rs:425-501) and its pieces with their definitions:
the packed red-mean key (range proof over all 2^24 pixel colours, argmin against the plain key incl. ties) and the
round-half-away of a target as two round-down additions.  The GPU parity tests remain the check of the kernel itself.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import binding as ob
from snesimage_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(tempfile.mkdtemp(prefix="dither_emu_"), "libdither_emu.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "dither_emulate.cpp")])
    lib = ctypes.CDLL(out)
    p = ctypes.c_void_p
    lib.dither_emulate.argtypes = [p, p, p, ctypes.c_int, ctypes.c_int, ctypes.c_int, p]
    lib.nearest_many.argtypes = [p, ctypes.c_int, p, ctypes.c_int, p]
    lib.round_many.argtypes = [p, ctypes.c_int, p]
    lib.key_constants.argtypes = [p]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def redmean_key(t, c):
    """common.cuh: redmean_key = 512 * color_distance_red_mean^2 (lib.rs:1080-1088), int64."""
    t = t.astype(np.int64)
    c = c.astype(np.int64)
    rs = t[..., 0] + c[..., 0]
    dr, dg, db = (t[..., k] - c[..., k] for k in range(3))
    return (1024 + rs) * dr * dr + 2048 * dg * dg + (1534 - rs) * db * db


def test_packed_key_fits_int32_for_every_pixel_colour(emu):
    """v = 8 key' + (j & 7) must be an int32 for every pixel colour and every entry colour: key' = key - T + L in [-2^28, 2^28)."""
    k = np.zeros(3, np.int64)
    emu.key_constants(_ptr(k))
    ls, lg, k0 = (int(v) for v in k)
    v = np.arange(256, dtype=np.int64)
    r, g, b = v[:, None, None], v[None, :, None], v[None, None, :]
    T = r ** 3 + 1024 * r * r + 2048 * g * g + 1534 * b * b - r * b * b
    D = ls * (r * r + b * b) + lg * g + k0 - T              # key' - key
    # largest key of pixel colour (r, g, b) over all entry colours: G and B at the far end, R by enumeration
    R = v[None, None, :]
    mb = np.maximum(v, 255 - v)[None, :, None]
    rr = v[:, None, None]
    krb = ((1024 + rr + R) * (R - rr) ** 2 + (1534 - rr - R) * mb * mb).max(axis=2)       # [r][b]
    kmax = krb[:, None, :] + (2048 * np.maximum(v, 255 - v) ** 2)[None, :, None]          # [r][g][b]
    assert D.min() >= -(2 ** 28)                 # key >= 0
    assert (kmax + D).max() < 2 ** 28
    assert 8 * (kmax + D).max() + 7 < 2 ** 31 and 8 * D.min() >= -(2 ** 31)


@pytest.mark.parametrize("S", [1, 2, 3, 7, 8, 9, 15, 16, 23, 64])
def test_nearest_rgb_is_the_first_minimum_of_the_plain_key(emu, S):
    rng = np.random.default_rng(100 + S)
    for trial in range(6):
        pal = rng.integers(0, 256, (S, 3)).astype(np.uint8)
        if trial >= 2 and S > 1:   # duplicated entries: ties must go to the lower index
            src = rng.integers(0, S, S)
            pal = pal[np.minimum(src, np.arange(S))]
        if trial == 5:
            pal[:] = rng.choice([0, 255], (S, 3)).astype(np.uint8)
        n = 20000
        tg = rng.integers(0, 256, (n, 3)).astype(np.uint8)
        tg[:64] = rng.choice([0, 255], (64, 3)).astype(np.uint8)
        tg[64:64 + S] = pal                                  # exact hits
        want = np.argmin(redmean_key(tg[:, None, :], pal[None, :, :]), axis=1).astype(np.int32)   # argmin = first minimum
        got = np.zeros(n, np.int32)
        emu.nearest_many(_ptr(pal), S, _ptr(tg), n, _ptr(got))
        assert np.array_equal(got, want)


def test_round_clamp_is_round_half_away_from_zero(emu):
    rng = np.random.default_rng(7)
    ints = np.arange(-300, 600, dtype=np.float64)
    ties = ints + 0.5
    cases = np.concatenate([
        ints, ties, np.nextafter(ties, -np.inf), np.nextafter(ties, np.inf), np.nextafter(ints, -np.inf), np.nextafter(ints, np.inf),
        np.array([0.49999999999999994, -0.49999999999999994, -0.5, -0.0, 0.0, 254.5, 254.49999999999997, 255.5, 1e9, -1e9]),
        rng.uniform(-400, 700, 200000), rng.integers(-10, 300, 50000) + rng.choice([0.5, 0.25, 0.75, 0.125], 50000)])
    got = np.zeros(len(cases), np.int32)
    emu.round_many(_ptr(np.ascontiguousarray(cases)), len(cases), _ptr(got))
    c = np.clip(cases, 0.0, 255.0)                  # lib.rs:773-778: clamp, then round() = half away from zero
    want = np.floor(c + 0.5)                        # c >= 0 and c + 0.5 is exact for c in [0, 255] except just below a tie ...
    frac = c - np.floor(c)                          # ... so decide by the (exact) fraction instead
    want = (np.floor(c) + (frac >= 0.5)).astype(np.int32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("family,C,S,seed", [("V", 8, 15, 1), ("T", 8, 15, 2), ("G", 4, 3, 3), ("B", 2, 4, 4), ("T", 3, 16, 5), ("V", 1, 7, 6)])
def test_wavefront_step_matches_oracle_optimize(emu, family, C, S, seed):
    rgba = synth.image(seed, family)
    o = ob.OracleImage(rgba, C, S, True, False, False)
    pal = synth.random_palette(seed, C, S)
    tp = synth.random_tile_palettes(seed, C)
    o.palette = pal
    o.tile_palettes = tp
    o.optimize()
    want = o.palette_map
    rgb8 = np.stack([ob.snes_as_rgba(c)[:3] for c in pal]).astype(np.uint8)
    got = np.full((256, 256), 77, np.uint8)
    emu.dither_emulate(_ptr(np.ascontiguousarray(rgba)), _ptr(tp), _ptr(np.ascontiguousarray(rgb8)), C, S, 0, _ptr(got))
    assert np.array_equal(got, want)
    # gi format: global entry number, 255 for a transparent pixel
    gi = np.zeros((256, 256), np.uint8)
    emu.dither_emulate(_ptr(np.ascontiguousarray(rgba)), _ptr(tp), _ptr(np.ascontiguousarray(rgb8)), C, S, 1, _ptr(gi))
    sub = np.repeat(np.repeat(tp.reshape(32, 32), 8, axis=0), 8, axis=1).astype(np.int64) * S
    assert np.array_equal(gi, np.where(rgba[..., 3] > 0, sub + want, 255).astype(np.uint8))
