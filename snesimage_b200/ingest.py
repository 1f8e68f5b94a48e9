"""Image ingest and JSON resume: the two data formats either side of the hot path.

* `load_rgba` is `image::open(path)?.into_rgba8()` of /root/reference/src/lib.rs:836 followed by the size check of
  lib.rs:838-840.  Host-side only (PIL decodes; nothing here touches the GPU).
* `state_from_json` inverts `OptimizedImage::as_json` (lib.rs:579-625), which the reference cannot do (TODO.md:38-39
  lists saving / resuming as a wish): the JSON document holds everything the optimiser state consists of except the
  schedule cursor.
"""
from __future__ import annotations

import json
from typing import Tuple, Union

import numpy as np

WIDTH = HEIGHT = 256          # lib.rs:29-30
TILES = 1024                  # lib.rs:58


def check_size(width: int, height: int):
    """lib.rs:838-840 rejects an image only if BOTH sides differ from 256 (`&&`); an image with exactly one side of
    256 then walks off the fixed 32x32 tile table (lib.rs:58, 565).  Both cases are refused here, the first with the
    reference's own message."""
    if width != WIDTH and height != HEIGHT:
        raise ValueError("Image size must be 256x256")
    if width != WIDTH or height != HEIGHT:
        raise ValueError(f"Image size must be 256x256 (got {width}x{height}; the reference lets this through and then "
                         f"indexes past its 32x32 tile table)")


def load_rgba(path: str) -> np.ndarray:
    """Decode any format PIL reads into RGBA8, (256, 256, 4) uint8, r,g,b,a byte order (the layout of `rgb::RGBA8`)."""
    from PIL import Image
    with Image.open(path) as im:
        check_size(im.width, im.height)
        rgba = np.asarray(im.convert("RGBA"), dtype=np.uint8)
    return np.ascontiguousarray(rgba)


def snes_color_from_u16(word: int) -> Tuple[int, int, int]:
    """Inverse of SnesColor::as_u16 (lib.rs:679-681): r | g << 5 | b << 10.  A component of 32 (the reference's
    `round(v / 8)` quirk, lib.rs:396-400) carries into the next field and cannot be recovered."""
    return word & 31, (word >> 5) & 31, (word >> 10) & 31


def state_from_json(doc: Union[str, dict], subpalette_count: int, subpalette_size: int):
    """(palette[C*S,3] u8, tile_palettes[1024] u8, palette_map[65536] u8, transparent[65536] bool) from the
    `{palette, tiles, tile_palettes}` document of lib.rs:579-625.

    palette: C rows of 16 words, slot 0 = 0, slots 1..=S = the colours (lib.rs:582-594).  tiles: 1024 arrays of 64
    values, tile row-major / in-tile row-major, palette_map + 1, or 0 for a transparent pixel (lib.rs:599-617)."""
    if isinstance(doc, str):
        doc = json.loads(doc)
    C, S = subpalette_count, subpalette_size
    pal_words = np.asarray(doc["palette"], dtype=np.int64).reshape(-1)
    if pal_words.size != 16 * C:
        raise ValueError(f"palette has {pal_words.size} words, expected {16 * C} for {C} subpalettes")
    if S > 15:
        raise ValueError("as_json stores at most 15 colours per subpalette")
    palette = np.zeros((C * S, 3), np.uint8)
    for p in range(C):
        for i in range(S):
            palette[p * S + i] = snes_color_from_u16(int(pal_words[16 * p + 1 + i]))
    tile_palettes = np.asarray(doc["tile_palettes"], dtype=np.int64).reshape(-1)
    if tile_palettes.size != TILES or tile_palettes.min() < 0 or tile_palettes.max() >= C:
        raise ValueError("tile_palettes must hold 1024 values below subpalette_count")
    tiles = np.asarray(doc["tiles"], dtype=np.int64)
    if tiles.shape != (TILES, 64) or tiles.min() < 0 or tiles.max() > S:
        raise ValueError("tiles must be 1024 arrays of 64 values in 0..=subpalette_size")
    # tile (ty, tx), in-tile (py, px) -> pixel (ty*8 + py, tx*8 + px)
    grid = tiles.reshape(32, 32, 8, 8).transpose(0, 2, 1, 3).reshape(HEIGHT * WIDTH)
    transparent = grid == 0
    palette_map = np.where(transparent, 0, grid - 1).astype(np.uint8)
    return palette, tile_palettes.astype(np.uint8), palette_map, transparent
