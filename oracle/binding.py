"""ctypes binding of the CPU oracle (oracle/libsnes_oracle.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under snesimage_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsnes_oracle.so")

NPIX = 65536
NTILES = 1024
TOTAL_SCALE_PIXELS = 87360
SCALE_DIMS = [256, 128, 64, 32, 16, 8]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("snes_oracle.c", "snes_oracle.h", "constants_unverified.h")]
    stale = force or not os.path.exists(_SO)
    if not stale and all(os.path.exists(s) for s in src):
        stale = any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsnes_oracle.so"])
    return _SO


_lib = None

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i, d, f = C.c_void_p, C.c_int, C.c_double, C.c_float
    sig = {
        "ora_image_new": (vp, [_u8p, i, i, i, i, i, i, i]),
        "ora_image_free": (None, [vp]),
        "ora_initialize_tiles": (i, [vp]),
        "ora_recalculate_palettes": (i, [vp]),
        "ora_optimize": (None, [vp]),
        "ora_error": (d, [vp]),
        "ora_as_rgba": (None, [vp, _u8p]),
        "ora_optimize_palette_entry_random": (i, [vp, i, i, _u8p, i]),
        "ora_optimize_palette_entry_nes": (i, [vp, i, i]),
        "ora_optimize_palette_entry_channel": (i, [vp, i, i, i]),
        "ora_eval_candidates": (None, [vp, i, i, _u8p, i, _f64p, vp]),
        "ora_get_palette": (None, [vp, _u8p]),
        "ora_set_palette": (None, [vp, _u8p]),
        "ora_get_tile_palettes": (None, [vp, _u8p]),
        "ora_set_tile_palettes": (None, [vp, _u8p]),
        "ora_get_palette_map": (None, [vp, _u8p]),
        "ora_set_palette_map": (None, [vp, _u8p]),
        "ora_as_json_arrays": (None, [vp, _u16p, _u8p, _u8p]),
        "ora_nes_color": (None, [i, _u8p]),
        "ora_snes_as_rgba": (None, [_u8p, _u8p]),
        "ora_snes_as_u16": (C.c_uint16, [_u8p]),
        "ora_new_nes_only": (None, [_u8p, i, _u8p]),
        "ora_color_distance_red_mean": (d, [_u8p, _u8p]),
        "ora_color_distance_cielab": (d, [_u8p, _u8p]),
        "ora_closest_color_index": (i, [_u8p, i, _f64p, i]),
        "ora_srgb8_to_lab_f32": (None, [C.c_uint8, C.c_uint8, C.c_uint8, _f32p]),
        "ora_lab_f64_to_srgb8": (None, [_f64p, _u8p]),
        "ora_ciede2000_f32": (f, [_f32p, _f32p]),
        "ora_ciede2000_f64": (d, [_f64p, _f64p]),
        "ora_srgb_eotf": (f, [f]),
        "ora_set_transfer_luts": (None, [vp, vp]),
        "ora_get_transfer_luts": (None, [_f32p, _f32p]),
        "ora_cbrtf": (f, [f]),
        "ora_linear_rgb_to_xyb": (None, [_f32p, _f32p]),
        "ora_gaussian_coeffs": (None, [_f32p, _f32p, C.POINTER(C.c_int)]),
        "ora_blur_plane": (None, [_f32p, _f32p, i, i]),
        "ora_kmeans": (i, [_f64p, i, i, _f64p, _i32p]),
        "ora_ssimulacra2_rgba8": (d, [_u8p, _u8p, i, i, vp]),
        "ora_xyb_pyramid_rgba8": (None, [_u8p, i, i, _f32p]),
        "ora_source_planes_rgba8": (None, [_u8p, i, i, _f32p, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _u8(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if shape is not None:
        a = a.reshape(shape)
    return a


class OracleImage:
    """struct OptimizedImage of the reference (lib.rs:33-626), CPU restatement."""

    def __init__(self, rgba, sub_count=1, sub_size=7, dither=False, perceptual_palettes=False, nes=False):
        rgba = _u8(rgba)
        assert rgba.shape == (256, 256, 4), rgba.shape
        self._l = lib()
        self.sub_count, self.sub_size = int(sub_count), int(sub_size)
        self.dither, self.perceptual_palettes, self.nes = bool(dither), bool(perceptual_palettes), bool(nes)
        self.rgba = rgba.copy()
        self._h = self._l.ora_image_new(self.rgba, 256, 256, self.sub_count, self.sub_size, int(self.dither),
                                        int(self.perceptual_palettes), int(self.nes))
        if not self._h:
            raise ValueError("oracle rejected the image/config")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ora_image_free(self._h)
            self._h = None

    # --- lib.rs methods -----------------------------------------------------------------------
    def initialize_tiles(self):
        if self._l.ora_initialize_tiles(self._h) != 0:
            raise RuntimeError("cogset Kmeans assertion (2 <= k < n) would fail")

    def recalculate_palettes(self):
        if self._l.ora_recalculate_palettes(self._h) != 0:
            raise RuntimeError("cogset Kmeans assertion (2 <= k < n) would fail")

    def optimize(self):
        self._l.ora_optimize(self._h)

    def error(self) -> float:
        return float(self._l.ora_error(self._h))

    def as_rgba(self):
        out = np.zeros((256, 256, 4), np.uint8)
        self._l.ora_as_rgba(self._h, out)
        return out

    def optimize_palette_entry_random(self, palette, index, cand):
        cand = _u8(cand, (-1, 3))
        self._l.ora_optimize_palette_entry_random(self._h, palette, index, cand, len(cand))

    def optimize_palette_entry_nes(self, palette, index):
        self._l.ora_optimize_palette_entry_nes(self._h, palette, index)

    def optimize_palette_entry_channel(self, palette, index, channel):
        self._l.ora_optimize_palette_entry_channel(self._h, palette, index, channel)

    def eval_candidates(self, palette, index, cand, want_maps=False):
        cand = _u8(cand, (-1, 3))
        scores = np.zeros(len(cand), np.float64)
        maps = np.zeros((len(cand), 256, 256), np.uint8) if want_maps else None
        self._l.ora_eval_candidates(self._h, palette, index, cand, len(cand), scores,
                                    maps.ctypes.data if want_maps else None)
        return (scores, maps) if want_maps else scores

    # --- state --------------------------------------------------------------------------------
    @property
    def palette(self):
        out = np.zeros((self.sub_count * self.sub_size, 3), np.uint8)
        self._l.ora_get_palette(self._h, out)
        return out

    @palette.setter
    def palette(self, v):
        self._l.ora_set_palette(self._h, _u8(v, (self.sub_count * self.sub_size, 3)))

    @property
    def tile_palettes(self):
        out = np.zeros(NTILES, np.uint8)
        self._l.ora_get_tile_palettes(self._h, out)
        return out

    @tile_palettes.setter
    def tile_palettes(self, v):
        self._l.ora_set_tile_palettes(self._h, _u8(v, (NTILES,)))

    @property
    def palette_map(self):
        out = np.zeros(NPIX, np.uint8)
        self._l.ora_get_palette_map(self._h, out)
        return out.reshape(256, 256)

    @palette_map.setter
    def palette_map(self, v):
        self._l.ora_set_palette_map(self._h, _u8(v, (NPIX,)))

    def as_json(self) -> dict:
        pal = np.zeros(self.sub_count * 16, np.uint16)
        tiles = np.zeros(NTILES * 64, np.uint8)
        tp = np.zeros(NTILES, np.uint8)
        self._l.ora_as_json_arrays(self._h, pal, tiles, tp)
        return {"palette": pal.tolist(), "tiles": tiles.reshape(NTILES, 64).tolist(), "tile_palettes": tp.tolist()}


# --- free functions ---------------------------------------------------------------------------
def ssimulacra2(src_rgba, dst_rgba, want_avg=False):
    src, dst = _u8(src_rgba), _u8(dst_rgba)
    h, w = src.shape[:2]
    avg = np.zeros((6, 18), np.float64)
    s = lib().ora_ssimulacra2_rgba8(src, dst, w, h, avg.ctypes.data if want_avg else None)
    return (float(s), avg) if want_avg else float(s)


def xyb_pyramid(rgba):
    out = np.zeros(TOTAL_SCALE_PIXELS * 3, np.float32)
    lib().ora_xyb_pyramid_rgba8(_u8(rgba), 256, 256, out)
    return out


def source_planes(rgba):
    mu1 = np.zeros(TOTAL_SCALE_PIXELS * 3, np.float32)
    s11 = np.zeros(TOTAL_SCALE_PIXELS * 3, np.float32)
    lib().ora_source_planes_rgba8(_u8(rgba), 256, 256, mu1, s11)
    return mu1, s11


def kmeans(points, k):
    pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    centres = np.zeros((k, 3), np.float64)
    assign = np.zeros(len(pts), np.int32)
    it = lib().ora_kmeans(pts, len(pts), k, centres, assign)
    return it, centres, assign


def set_transfer_luts(yuvxyb_eotf=None, palette_eotf=None):
    """Swap in 256-entry sRGB -> linear tables dumped from the real crates (None: the built-in one, libm powf)."""
    a = None if yuvxyb_eotf is None else np.ascontiguousarray(yuvxyb_eotf, np.float32).reshape(256)
    b = None if palette_eotf is None else np.ascontiguousarray(palette_eotf, np.float32).reshape(256)
    lib().ora_set_transfer_luts(None if a is None else a.ctypes.data, None if b is None else b.ctypes.data)


def transfer_luts():
    a, b = np.zeros(256, np.float32), np.zeros(256, np.float32)
    lib().ora_get_transfer_luts(a, b)
    return a, b


def ciede2000_f64(l1, l2):
    return float(lib().ora_ciede2000_f64(np.asarray(l1, np.float64), np.asarray(l2, np.float64)))


def ciede2000_f32(l1, l2):
    return float(lib().ora_ciede2000_f32(np.asarray(l1, np.float32), np.asarray(l2, np.float32)))


def srgb8_to_lab(r, g, b):
    out = np.zeros(3, np.float32)
    lib().ora_srgb8_to_lab_f32(int(r), int(g), int(b), out)
    return out


def lab_to_srgb8(lab):
    out = np.zeros(3, np.uint8)
    lib().ora_lab_f64_to_srgb8(np.asarray(lab, np.float64), out)
    return out


def red_mean(a, b):
    return float(lib().ora_color_distance_red_mean(_u8(a), _u8(b)))


def cielab(a, b):
    return float(lib().ora_color_distance_cielab(_u8(a), _u8(b)))


def nes_color(i):
    out = np.zeros(3, np.uint8)
    lib().ora_nes_color(int(i), out)
    return out


def new_nes_only(c5, cielab_flag=False):
    out = np.zeros(3, np.uint8)
    lib().ora_new_nes_only(_u8(c5), int(cielab_flag), out)
    return out


def snes_as_rgba(c5):
    out = np.zeros(4, np.uint8)
    lib().ora_snes_as_rgba(_u8(c5), out)
    return out


def closest_color_index(colors5, target, cielab_flag=False):
    colors5 = _u8(colors5, (-1, 3))
    return int(lib().ora_closest_color_index(colors5, len(colors5), np.asarray(target, np.float64), int(cielab_flag)))


def gaussian_coeffs():
    n2, d1 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    r = C.c_int(0)
    lib().ora_gaussian_coeffs(n2, d1, C.byref(r))
    return n2, d1, r.value


def blur_plane(plane):
    plane = np.ascontiguousarray(plane, np.float32)
    out = np.zeros_like(plane)
    lib().ora_blur_plane(plane, out, plane.shape[1], plane.shape[0])
    return out
