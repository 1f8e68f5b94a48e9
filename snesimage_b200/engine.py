"""Host-side mirror of the reference's `OptimizedImage` (/root/reference/src/lib.rs:33-626) over the
C ABI of libsnesgpu.so (include/snesgpu.h).

Same method names, argument meaning and error behaviour as the reference's private type, so parity
tests read like tests of lib.rs: `initialize_tiles`, `recalculate_palettes`, `optimize`, `error`,
`optimize_palette_entry_{random,nes,channel}`, `as_rgba`, `as_json`.  The one deliberate change is
that `optimize_palette_entry_random` takes its trial colours as an explicit list: the reference
draws them from an unseeded `rand::rng()` (lib.rs:201-208), which no test could reproduce.

There is no CPU fallback.  If libsnesgpu.so cannot be loaded, or no sm_100 GPU is present, the calls
raise; nothing here imports or calls the oracle.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _build

NPIX = 65536
NTILES = 1024
TOTAL_SCALE_PIXELS = 87360
NES_COLOR_COUNT = 56

SNES_OK, SNES_E_INVALID, SNES_E_CUDA, SNES_E_KMEANS, SNES_E_NOMEM = 0, -1, -2, -3, -4


class SnesGpuError(RuntimeError):
    def __init__(self, code: int, message: str, context: str = ""):
        self.code = code
        super().__init__(f"{context}: {message}" if context else message)


class KmeansAssertion(SnesGpuError):
    """cogset's `assert!(2 <= k && k < data.len())` would panic in the reference."""


class Best(C.Structure):
    _fields_ = [("err", C.c_double), ("idx", C.c_int32), ("pad", C.c_int32)]


BEST_DTYPE = np.dtype([("err", "<f8"), ("idx", "<i4"), ("pad", "<i4")])


class Step(C.Structure):
    """snes_step: the palette entry (and channel) one iteration of run() works on."""
    _fields_ = [("palette", C.c_int32), ("index", C.c_int32), ("channel", C.c_int32), ("reserved", C.c_int32)]


class _Config(C.Structure):
    _fields_ = [("subpalette_count", C.c_int32), ("subpalette_size", C.c_int32), ("dither", C.c_uint8),
                ("perceptual_palettes", C.c_uint8), ("nes", C.c_uint8), ("reserved", C.c_uint8)]


@dataclass
class Config:
    """config::Config (/root/reference/src/config.rs:3-31); same field names and defaults."""
    source_filename: str = ""
    target_filename: str = ""
    subpalette_count: int = 1
    subpalette_size: int = 7
    dither: bool = False
    perceptual_palettes: bool = False
    nes: bool = False

    def _c(self) -> _Config:
        return _Config(int(self.subpalette_count), int(self.subpalette_size), int(bool(self.dither)),
                       int(bool(self.perceptual_palettes)), int(bool(self.nes)), 0)


_lib = None

# every symbol include/snesgpu.h declares: name -> (restype, argtypes)
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_SIGNATURES = {
    "snes_last_error": (C.c_char_p, []),
    "snes_version": (_i, []),
    "snes_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "snes_ctx_destroy": (None, [_vp]),
    "snes_ctx_kernel_launches": (C.c_int64, [_vp]),
    "snes_ctx_set_stream": (_i, [_vp, _vp]),
    "snes_ctx_synchronize": (_i, [_vp]),
    "snes_ctx_set_chunk": (_i, [_vp, _i]),
    "snes_ctx_set_all_terms": (_i, [_vp, _i]),
    "snes_ctx_cbrt_selfcheck": (_i, [_vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "snes_ctx_set_transfer_luts": (_i, [_vp, _vp, _vp]),
    "snes_ctx_set_scorer": (_i, [_vp, _i, _i, _i]),
    "snes_ctx_profile_begin": (_i, [_vp]),
    "snes_ctx_profile_only": (_i, [_vp, C.c_char_p]),
    "snes_ctx_profile_end": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "snes_image_new": (_i, [_vp, _vp, _i, _i, C.POINTER(_Config), C.POINTER(_vp)]),
    "snes_image_free": (None, [_vp]),
    "snes_image_initialize_tiles": (_i, [_vp]),
    "snes_image_recalculate_palettes": (_i, [_vp]),
    "snes_image_optimize": (_i, [_vp]),
    "snes_image_error": (_i, [_vp, C.POINTER(C.c_double)]),
    "snes_image_as_rgba": (_i, [_vp, _vp]),
    "snes_image_as_json": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "snes_image_optimize_palette_entry_random": (_i, [_vp, _i, _i, _vp, _i]),
    "snes_image_optimize_palette_entry_nes": (_i, [_vp, _i, _i]),
    "snes_image_optimize_palette_entry_channel": (_i, [_vp, _i, _i, _i]),
    "snes_image_get_palette": (_i, [_vp, _vp]),
    "snes_image_set_palette": (_i, [_vp, _vp]),
    "snes_image_get_tile_palettes": (_i, [_vp, _vp]),
    "snes_image_set_tile_palettes": (_i, [_vp, _vp]),
    "snes_image_get_palette_map": (_i, [_vp, _vp]),
    "snes_image_set_palette_map": (_i, [_vp, _vp]),
    "snes_batch_initialize_tiles": (_i, [_vp, _vp, _i]),
    "snes_batch_recalculate_palettes": (_i, [_vp, _vp, _i]),
    "snes_batch_optimize": (_i, [_vp, _vp, _i]),
    "snes_batch_error": (_i, [_vp, _vp, _i, _vp]),
    "snes_batch_error_dev": (_i, [_vp, _vp, _i, _vp]),
    "snes_batch_eval_candidates": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "snes_batch_eval_candidates_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "snes_batch_error_eval_candidates_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "snes_batch_error_eval_candidates_slice_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "snes_batch_step_random_shard_begin": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    "snes_batch_step_random_shard_end": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "snes_image_state_checksum": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "snes_comm_unique_id": (_i, [_vp]),
    "snes_ctx_comm_init": (_i, [_vp, _vp, _i, _i]),
    "snes_ctx_comm_destroy": (_i, [_vp]),
    "snes_dist_plan": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "snes_dist_step_random": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "snes_batch_eval_candidates_multi": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "snes_image_iterate": (_i, [_vp, _i, _vp, _i, _vp, _i, C.POINTER(_i), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "snes_batch_apply_best_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "snes_merge_best_dev": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "snes_batch_step_random": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "snes_batch_step_nes": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "snes_batch_step_channel": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "snes_batch_eval_tile_moves": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "snes_batch_step_tile_moves": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "snes_closest_color_index": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "snes_new_nes_only": (_i, [_vp, _vp, _i, _i, _vp]),
    "snes_image_debug_planes": (_i, [_vp, _vp, _vp, _vp]),
    "snes_image_debug_lab": (_i, [_vp, _vp]),
    "snes_image_kmeans_debug": (_i, [_vp, _vp, _vp]),
}


def library_path() -> str:
    return _build.SO


def lib():
    """Load libsnesgpu.so, (re)building it in-tree first when it is missing or older than a source it is built from
    (a no-op when fresh; on a box without nvcc a stale library is an error, not a silent old binary)."""
    global _lib
    if _lib is not None:
        return _lib
    so = os.environ.get("SNESGPU_SO") or _build.SO   # SNESGPU_SO: an instrumented debug build (scripts/phase_timing.py)
    if so == _build.SO:
        _build.build_library()
    L = C.CDLL(so)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc: int, context: str = ""):
    if rc == SNES_OK:
        return
    msg = lib().snes_last_error().decode("utf-8", "replace")
    if rc == SNES_E_KMEANS:
        raise KmeansAssertion(rc, msg, context)
    raise SnesGpuError(rc, msg, context)


def _u8(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(shape) if shape is not None else a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Context:
    """One per GPU (one process per GPU).  Owns the stream, scratch memory and the BGR555->Lab table."""

    def __init__(self, device: int = 0, chunk: Optional[int] = None):
        self._l = lib()
        h = _vp()
        _check(self._l.snes_ctx_create(int(device), C.byref(h)), "snes_ctx_create")
        self._h = h
        self.device = int(device)
        if chunk:
            self.set_chunk(chunk)

    def close(self):
        if getattr(self, "_h", None):
            self._l.snes_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(self._l.snes_ctx_kernel_launches(self._h))

    def set_stream(self, cuda_stream: Optional[int]):
        _check(self._l.snes_ctx_set_stream(self._h, cuda_stream), "snes_ctx_set_stream")

    def set_chunk(self, evaluations: int):
        _check(self._l.snes_ctx_set_chunk(self._h, int(evaluations)), "snes_ctx_set_chunk")

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """The library's own NCCL communicator (collective: every rank calls it with the id rank 0 got from comm_unique_id)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self._l.snes_ctx_comm_init(self._h, buf, int(rank), int(world)), "snes_ctx_comm_init")

    def comm_destroy(self):
        _check(self._l.snes_ctx_comm_destroy(self._h), "snes_ctx_comm_destroy")

    def set_transfer_luts(self, yuvxyb_eotf=None, palette_eotf=None):
        """Swap in verified 256-entry sRGB -> linear tables (None: the built-in one).  Before any image is created."""
        a = None if yuvxyb_eotf is None else np.ascontiguousarray(yuvxyb_eotf, np.float32).reshape(256)
        b = None if palette_eotf is None else np.ascontiguousarray(palette_eotf, np.float32).reshape(256)
        _check(self._l.snes_ctx_set_transfer_luts(self._h, _ptr(a), _ptr(b)), "snes_ctx_set_transfer_luts")

    def set_all_terms(self, on: bool):
        """on: the scorer computes ssim_map even where its pooling weights are zero (A/B check of the edge-only pair items)."""
        _check(self._l.snes_ctx_set_all_terms(self._h, int(bool(on))), "snes_ctx_set_all_terms")

    def cbrt_selfcheck(self, lo: float, hi: float):
        """Compare the kernels' cube root with the restated msun cbrtf on every float in [lo, hi): (mismatches, fallbacks)."""
        lo_bits, hi_bits = (int(np.float32(v).view(np.uint32)) for v in (lo, hi))
        bad, fb = C.c_uint64(0), C.c_uint64(0)
        _check(self._l.snes_ctx_cbrt_selfcheck(self._h, lo_bits, hi_bits, C.byref(bad), C.byref(fb)), "snes_ctx_cbrt_selfcheck")
        return bad.value, fb.value

    def set_scorer(self, fused: int = 3, block_width: int = 32, delta_assign: bool = True):
        """fused: 3 = k_score_v3 (default), 2 = k_score_v2 (its predecessor, kept as the A/B check)."""
        _check(self._l.snes_ctx_set_scorer(self._h, int(fused), int(block_width), int(delta_assign)), "snes_ctx_set_scorer")

    def profile_begin(self, only: Optional[str] = None):
        """Start per-launch CUDA-event timing; `only`: bracket just the launches whose kernel name contains it."""
        _check(self._l.snes_ctx_profile_only(self._h, only.encode() if only else None), "snes_ctx_profile_only")
        _check(self._l.snes_ctx_profile_begin(self._h), "snes_ctx_profile_begin")

    def profile_end(self) -> dict:
        """Per-kernel device time since profile_begin(): {kernel: {"ms": total, "n": launches}}."""
        buf = C.create_string_buffer(1 << 16)
        n = _sz(0)
        _check(self._l.snes_ctx_profile_end(self._h, buf, len(buf), C.byref(n)), "snes_ctx_profile_end")
        return json.loads(buf.value.decode("ascii"))

    def synchronize(self):
        _check(self._l.snes_ctx_synchronize(self._h), "snes_ctx_synchronize")

    # ---- colour primitives (lib.rs:628-795) ----------------------------------------------------
    def closest_color_index(self, colors5, targets, cielab: bool = False) -> np.ndarray:
        colors5 = _u8(colors5, (-1, 3))
        targets = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
        out = np.zeros(len(targets), np.int32)
        _check(self._l.snes_closest_color_index(self._h, _ptr(colors5), len(colors5), _ptr(targets), len(targets),
                                                int(cielab), _ptr(out)), "snes_closest_color_index")
        return out

    def new_nes_only(self, colors5, cielab: bool = False) -> np.ndarray:
        colors5 = _u8(colors5, (-1, 3))
        out = np.zeros_like(colors5)
        _check(self._l.snes_new_nes_only(self._h, _ptr(colors5), len(colors5), int(cielab), _ptr(out)), "snes_new_nes_only")
        return out


def _handles(images: Sequence["OptimizedImage"]):
    arr = (_vp * len(images))(*[im._h for im in images])
    return arr


class OptimizedImage:
    """struct OptimizedImage (lib.rs:33-626) with its state resident on the GPU."""

    def __init__(self, ctx: Context, rgba, config: Config):
        rgba = _u8(rgba)
        if rgba.ndim != 3 or rgba.shape[2] != 4:
            raise ValueError("rgba must be an (H, W, 4) uint8 array")
        self.ctx = ctx
        self.config = config
        self._l = ctx._l
        h = _vp()
        cfg = config._c()
        _check(self._l.snes_image_new(ctx._h, _ptr(rgba), rgba.shape[1], rgba.shape[0], C.byref(cfg), C.byref(h)),
               "OptimizedImage::new")
        self._h = h
        self.sub_count, self.sub_size = int(config.subpalette_count), int(config.subpalette_size)

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._l.snes_image_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- lib.rs methods --------------------------------------------------------------------------
    def initialize_tiles(self):
        _check(self._l.snes_image_initialize_tiles(self._h), "Unable to initialize tiles")

    def recalculate_palettes(self):
        _check(self._l.snes_image_recalculate_palettes(self._h), "Unable to recalculate palettes")

    def optimize(self):
        _check(self._l.snes_image_optimize(self._h), "Unable to optimize image")

    def error(self) -> float:
        v = C.c_double(0.0)
        _check(self._l.snes_image_error(self._h, C.byref(v)), "Failed to compute SSIMULACRA2")
        return float(v.value)

    def as_rgba(self) -> np.ndarray:
        out = np.zeros((256, 256, 4), np.uint8)
        _check(self._l.snes_image_as_rgba(self._h, _ptr(out)), "as_rgba")
        return out

    def optimize_palette_entry_random(self, palette: int, index: int, cand):
        cand = _u8(cand, (-1, 3))
        _check(self._l.snes_image_optimize_palette_entry_random(self._h, palette, index, _ptr(cand), len(cand)),
               "Unable to optimize palette with the random method")

    def optimize_palette_entry_nes(self, palette: int, index: int):
        _check(self._l.snes_image_optimize_palette_entry_nes(self._h, palette, index),
               "Unable to optimize palette with the NES method")

    def optimize_palette_entry_channel(self, palette: int, index: int, channel: int):
        _check(self._l.snes_image_optimize_palette_entry_channel(self._h, palette, index, channel),
               "Unable to optimize palette with the channel method")

    def iterate(self, mode: str, steps, cand=None):
        """`len(steps)` consecutive iterations of run()'s loop (lib.rs:889-910) in one call, speculatively (see
        snes_image_iterate).  steps: (palette, index[, channel]) tuples; cand (random mode): (len(steps), ncand, 3).
        Returns (iterations consumed, error() before, error() of the new state); error before is NaN in NES mode."""
        arr, n = _steps(steps)
        m = {"random": 0, "nes": 1, "channel": 2}[mode]
        ncand = 0
        if m == 0:
            cand = _u8(cand).reshape(n, -1, 3)
            ncand = cand.shape[1]
        used, before, err = _i(0), C.c_double(float("nan")), C.c_double(0.0)
        _check(self._l.snes_image_iterate(self._h, m, arr, n, _ptr(cand) if m == 0 else None, ncand, C.byref(used), C.byref(before),
                                          C.byref(err)), "Unable to optimize palette")
        return int(used.value), float(before.value), float(err.value)

    def as_json_string(self) -> str:
        n = _sz(0)
        _check(self._l.snes_image_as_json(self._h, None, 0, C.byref(n)), "as_json")
        buf = C.create_string_buffer(n.value + 1)
        _check(self._l.snes_image_as_json(self._h, buf, n.value + 1, C.byref(n)), "as_json")
        return buf.value.decode("ascii")

    def as_json(self) -> dict:
        return json.loads(self.as_json_string())

    # ---- candidate loops (lib.rs:205-220, 252-262, 296-306) ----------------------------------------
    def eval_candidates(self, palette: int, index: int, cand, want_maps: bool = False):
        r = batch_eval_candidates([self], palette, index, _u8(cand, (1, -1, 3)), want_maps=want_maps, want_best=False)
        return (r["scores"][0], r["maps"][0]) if want_maps else r["scores"][0]

    # ---- state -------------------------------------------------------------------------------------
    @property
    def palette(self) -> np.ndarray:
        out = np.zeros((self.sub_count * self.sub_size, 3), np.uint8)
        _check(self._l.snes_image_get_palette(self._h, _ptr(out)), "get_palette")
        return out

    @palette.setter
    def palette(self, v):
        v = _u8(v, (self.sub_count * self.sub_size, 3))
        _check(self._l.snes_image_set_palette(self._h, _ptr(v)), "set_palette")

    @property
    def tile_palettes(self) -> np.ndarray:
        out = np.zeros(NTILES, np.uint8)
        _check(self._l.snes_image_get_tile_palettes(self._h, _ptr(out)), "get_tile_palettes")
        return out

    @tile_palettes.setter
    def tile_palettes(self, v):
        v = _u8(v, (NTILES,))
        _check(self._l.snes_image_set_tile_palettes(self._h, _ptr(v)), "set_tile_palettes")

    @property
    def palette_map(self) -> np.ndarray:
        out = np.zeros(NPIX, np.uint8)
        _check(self._l.snes_image_get_palette_map(self._h, _ptr(out)), "get_palette_map")
        return out.reshape(256, 256)

    @palette_map.setter
    def palette_map(self, v):
        v = _u8(v, (NPIX,))
        _check(self._l.snes_image_set_palette_map(self._h, _ptr(v)), "set_palette_map")

    def state_checksum(self) -> int:
        """64-bit FNV-1a over palette, tile_palettes and palette_map."""
        v = C.c_uint64(0)
        _check(self._l.snes_image_state_checksum(self._h, C.byref(v)), "state_checksum")
        return int(v.value)

    # ---- debug taps --------------------------------------------------------------------------------
    def debug_planes(self):
        xyb = np.zeros(TOTAL_SCALE_PIXELS * 3, np.float32)
        mu1 = np.zeros_like(xyb)
        s11 = np.zeros_like(xyb)
        _check(self._l.snes_image_debug_planes(self._h, _ptr(xyb), _ptr(mu1), _ptr(s11)), "debug_planes")
        return xyb, mu1, s11

    def debug_lab(self) -> np.ndarray:
        out = np.zeros((NPIX, 4), np.float32)
        _check(self._l.snes_image_debug_lab(self._h, _ptr(out)), "debug_lab")
        return out[:, :3]

    def kmeans_debug(self):
        centres = np.zeros((256, 3), np.float64)
        status = np.zeros(512, np.int32)
        _check(self._l.snes_image_kmeans_debug(self._h, _ptr(centres), _ptr(status)), "kmeans_debug")
        return centres, status


# ---- batch API: many independent OptimizedImages sharing one Config ------------------------------
def _ctx_of(images: Sequence[OptimizedImage]) -> Context:
    if not images:
        raise ValueError("empty batch")
    return images[0].ctx


def batch_initialize_tiles(images: Sequence[OptimizedImage]):
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_initialize_tiles(ctx._h, _handles(images), len(images)), "Unable to initialize tiles")


def batch_recalculate_palettes(images: Sequence[OptimizedImage]):
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_recalculate_palettes(ctx._h, _handles(images), len(images)), "Unable to recalculate palettes")


def batch_optimize(images: Sequence[OptimizedImage]):
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_optimize(ctx._h, _handles(images), len(images)), "Unable to optimize image")


def batch_error(images: Sequence[OptimizedImage]) -> np.ndarray:
    ctx = _ctx_of(images)
    out = np.zeros(len(images), np.float64)
    _check(ctx._l.snes_batch_error(ctx._h, _handles(images), len(images), _ptr(out)), "Failed to compute SSIMULACRA2")
    return out


def batch_eval_candidates(images: Sequence[OptimizedImage], palette: int, index: int, cand, want_scores: bool = True,
                          want_maps: bool = False, want_best: bool = True) -> dict:
    """For every image j and candidate k: entry (palette, index) := cand[j][k]; optimize(); error().
    cand: (nimg, ncand, 3) 5-bit colours.  The images' own state is not modified."""
    ctx = _ctx_of(images)
    nimg = len(images)
    cand = _u8(cand).reshape(nimg, -1, 3)
    ncand = cand.shape[1]
    scores = np.zeros((nimg, ncand), np.float64) if want_scores else None
    maps = np.zeros((nimg, ncand, 256, 256), np.uint8) if want_maps else None
    best = np.zeros(nimg, BEST_DTYPE) if want_best else None
    _check(ctx._l.snes_batch_eval_candidates(ctx._h, _handles(images), nimg, palette, index, _ptr(cand), ncand, _ptr(scores),
                                             _ptr(maps), _ptr(best)), "snes_batch_eval_candidates")
    return {"scores": scores, "maps": maps, "best": best}


def _steps(steps):
    steps = [tuple(s) + (0,) * (3 - len(tuple(s))) for s in steps]
    arr = (Step * len(steps))(*[Step(int(p), int(i), int(ch), 0) for p, i, ch in steps])
    return arr, len(steps)


def batch_eval_candidates_multi(images: Sequence[OptimizedImage], steps, cand) -> dict:
    """Candidates of several palette entries against ONE state in one launch sequence.  steps: (palette, index) tuples;
    cand: (nimg, nsteps, ncand, 3).  Returns scores (nimg, nsteps, ncand) and best (nimg, nsteps)."""
    ctx = _ctx_of(images)
    nimg = len(images)
    arr, n = _steps(steps)
    cand = _u8(cand).reshape(nimg, n, -1, 3)
    ncand = cand.shape[2]
    scores = np.zeros((nimg, n, ncand), np.float64)
    best = np.zeros((nimg, n), BEST_DTYPE)
    _check(ctx._l.snes_batch_eval_candidates_multi(ctx._h, _handles(images), nimg, arr, n, _ptr(cand), ncand, _ptr(scores), _ptr(best)),
           "snes_batch_eval_candidates_multi")
    return {"scores": scores, "best": best}


def batch_step_random(images: Sequence[OptimizedImage], palette: int, index: int, cand, want_errors: bool = False):
    """optimize_palette_entry_random (lib.rs:191-240) on every image; cand: (nimg, ncand, 3)."""
    ctx = _ctx_of(images)
    nimg = len(images)
    cand = _u8(cand).reshape(nimg, -1, 3)
    best = np.zeros(nimg, BEST_DTYPE)
    errs = np.zeros(nimg, np.float64) if want_errors else None
    _check(ctx._l.snes_batch_step_random(ctx._h, _handles(images), nimg, palette, index, _ptr(cand), cand.shape[1], _ptr(best),
                                         _ptr(errs)), "Unable to optimize palette with the random method")
    return best, errs


def batch_step_nes(images: Sequence[OptimizedImage], palette: int, index: int, want_errors: bool = False):
    ctx = _ctx_of(images)
    nimg = len(images)
    best = np.zeros(nimg, BEST_DTYPE)
    errs = np.zeros(nimg, np.float64) if want_errors else None
    _check(ctx._l.snes_batch_step_nes(ctx._h, _handles(images), nimg, palette, index, _ptr(best), _ptr(errs)),
           "Unable to optimize palette with the NES method")
    return best, errs


def batch_step_channel(images: Sequence[OptimizedImage], palette: int, index: int, channel: int, want_errors: bool = False):
    ctx = _ctx_of(images)
    nimg = len(images)
    best = np.zeros(nimg, BEST_DTYPE)
    errs = np.zeros(nimg, np.float64) if want_errors else None
    _check(ctx._l.snes_batch_step_channel(ctx._h, _handles(images), nimg, palette, index, channel, _ptr(best), _ptr(errs)),
           "Unable to optimize palette with the channel method")
    return best, errs


# ---- device-pointer (asynchronous) entry points, used by bench.py and the multi-GPU driver --------
def _moves(moves, nimg: int) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(moves, dtype=np.int32))
    return m.reshape(nimg, -1, 2)


def batch_eval_tile_moves(images: Sequence[OptimizedImage], moves, want_maps: bool = False) -> dict:
    """Tile reassignment as evaluated candidates (TODO.md:36-37): for image j and move k = (tile, subpalette),
    tile_palettes[tile] = subpalette; optimize(); error().  moves: (nimg, nmoves, 2) int32, tile = tile_y*32 + tile_x.
    The images' own state is not modified."""
    ctx = _ctx_of(images)
    nimg = len(images)
    m = _moves(moves, nimg)
    nmoves = m.shape[1]
    scores = np.zeros((nimg, nmoves), np.float64)
    maps = np.zeros((nimg, nmoves, 256, 256), np.uint8) if want_maps else None
    best = np.zeros(nimg, BEST_DTYPE)
    _check(ctx._l.snes_batch_eval_tile_moves(ctx._h, _handles(images), nimg, _ptr(m), nmoves, _ptr(scores), _ptr(maps), _ptr(best)),
           "snes_batch_eval_tile_moves")
    return {"scores": scores, "maps": maps, "best": best}


def batch_step_tile_moves(images: Sequence[OptimizedImage], moves) -> dict:
    """Evaluate the moves and apply each image's best one if it is strictly better than the image's error()."""
    ctx = _ctx_of(images)
    nimg = len(images)
    m = _moves(moves, nimg)
    best = np.zeros(nimg, BEST_DTYPE)
    applied = np.zeros(nimg, np.uint8)
    _check(ctx._l.snes_batch_step_tile_moves(ctx._h, _handles(images), nimg, _ptr(m), m.shape[1], _ptr(best), _ptr(applied)),
           "snes_batch_step_tile_moves")
    return {"best": best, "applied": applied.astype(bool)}


def batch_error_dev(images: Sequence[OptimizedImage], d_errors: Optional[int] = None):
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_error_dev(ctx._h, _handles(images), len(images), d_errors), "snes_batch_error_dev")


def batch_eval_candidates_dev(images: Sequence[OptimizedImage], palette: int, index: int, d_cand: int, ncand: int,
                              cand_idx_base: int = 0, d_scores: Optional[int] = None, d_best: Optional[int] = None):
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_eval_candidates_dev(ctx._h, _handles(images), len(images), palette, index, d_cand, ncand,
                                                 cand_idx_base, d_scores, d_best), "snes_batch_eval_candidates_dev")


def batch_error_eval_candidates_dev(images: Sequence[OptimizedImage], palette: int, index: int, d_cand: int, ncand: int,
                                    cand_idx_base: int = 0, d_scores: Optional[int] = None, d_best: Optional[int] = None):
    """`best_error = self.error()` (lib.rs:199, 294) and the candidate loop in one pass: the images' own scoring items
    share the candidates' scorer launch."""
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_error_eval_candidates_dev(ctx._h, _handles(images), len(images), palette, index, d_cand, ncand,
                                                       cand_idx_base, d_scores, d_best), "snes_batch_error_eval_candidates_dev")


def batch_error_eval_candidates_slice_dev(images: Sequence[OptimizedImage], palette: int, index: int, d_cand_all: int, ncand_all: int,
                                          cand_lo: int, ncand: int, d_scores: Optional[int] = None, d_best: Optional[int] = None):
    """A rank's share of a candidate-sharded step: error() + candidates [cand_lo, cand_lo + ncand) of the full device list,
    read in place; best records carry indices into the full list."""
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_error_eval_candidates_slice_dev(ctx._h, _handles(images), len(images), palette, index, d_cand_all,
                                                             ncand_all, cand_lo, ncand, d_scores, d_best),
           "snes_batch_error_eval_candidates_slice_dev")


def batch_step_random_shard_begin(images: Sequence[OptimizedImage], palette: int, index: int, cand, cand_lo: int, ncand: int,
                                  d_best_local: int):
    """First half of a sharded step with host buffers: cand (nimg, ncand_all, 3) host array -> device, error() + this
    rank's slice, records left in d_best_local (device)."""
    ctx = _ctx_of(images)
    nimg = len(images)
    cand = _u8(cand).reshape(nimg, -1, 3)
    _check(ctx._l.snes_batch_step_random_shard_begin(ctx._h, _handles(images), nimg, palette, index, _ptr(cand), cand.shape[1],
                                                     cand_lo, ncand, d_best_local), "snes_batch_step_random_shard_begin")


def batch_step_random_shard_end(images: Sequence[OptimizedImage], palette: int, index: int, d_gathered: int, nranks: int,
                                rank_stride: int, best: Optional[np.ndarray] = None, want_errors: bool = False):
    """Second half: merge the gathered records of the ranks sharing these images, accept, optimize(); synchronous."""
    ctx = _ctx_of(images)
    nimg = len(images)
    if best is None:
        best = np.zeros(nimg, BEST_DTYPE)
    errs = np.zeros(nimg, np.float64) if want_errors else None
    _check(ctx._l.snes_batch_step_random_shard_end(ctx._h, _handles(images), nimg, palette, index, d_gathered, nranks, rank_stride,
                                                   _ptr(best), _ptr(errs)), "snes_batch_step_random_shard_end")
    return best, errs


def comm_unique_id() -> bytes:
    """A 128-byte NCCL unique id (call on one rank, hand the bytes to the others)."""
    buf = (C.c_uint8 * 128)()
    _check(lib().snes_comm_unique_id(buf), "snes_comm_unique_id")
    return bytes(buf)


def dist_plan(nimg_total: int, rank: int, world: int) -> dict:
    """snes_dist_plan: the library's own statement of driver.plan_shards."""
    v = [_i(0) for _ in range(5)]
    _check(lib().snes_dist_plan(int(nimg_total), int(rank), int(world), *[C.byref(x) for x in v]), "snes_dist_plan")
    return dict(zip(("img_lo", "img_hi", "cand_ranks", "slice", "slots"), (int(x.value) for x in v)))


def dist_step_random(images: Sequence[OptimizedImage], nimg_total: int, palette: int, index: int, cand, world_slots: int = 0):
    """One sharded optimize_palette_entry_random with the all-gather inside the library (snes_dist_step_random).
    cand: (local images, ncand, 3).  Returns (winning records of the local images, every rank's records or None)."""
    ctx = _ctx_of(images)
    nimg = len(images)
    cand = _u8(cand).reshape(nimg, -1, 3)
    best = np.zeros(nimg, BEST_DTYPE)
    allb = np.zeros(world_slots, BEST_DTYPE) if world_slots else None
    _check(ctx._l.snes_dist_step_random(ctx._h, _handles(images), nimg, int(nimg_total), palette, index, _ptr(cand), cand.shape[1],
                                        _ptr(best), _ptr(allb)), "snes_dist_step_random")
    return best, allb


def batch_apply_best_dev(images: Sequence[OptimizedImage], palette: int, index: int, d_cand_all: int, ncand_all: int, d_best: int):
    ctx = _ctx_of(images)
    _check(ctx._l.snes_batch_apply_best_dev(ctx._h, _handles(images), len(images), palette, index, d_cand_all, ncand_all, d_best),
           "snes_batch_apply_best_dev")


def merge_best_dev(ctx: Context, d_gathered: int, nranks: int, nimg: int, d_out: int, rank_stride: Optional[int] = None):
    """d_gathered: nranks blocks of records, rank r's at d_gathered + r * rank_stride * 16 (default stride: nimg)."""
    _check(ctx._l.snes_merge_best_dev(ctx._h, d_gathered, nranks, rank_stride or nimg, nimg, d_out), "snes_merge_best_dev")
