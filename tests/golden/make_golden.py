"""Generates tests/golden/oracle_vectors.json from the CPU oracle (oracle/).

The reference ships no golden vectors and cannot be built here (Rust crate, un-vendored dependencies),
so these are regression pins of the oracle restatement on seeded synthetic inputs, not
reference-derived values.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402
from snesimage_b200 import synth  # noqa: E402


def _sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute() -> dict:
    out = {}
    for fam in "VGBT":
        out[f"image_{fam}0_sha256"] = _sha(synth.image(0, fam))
    out["candidates_0_0_sha256"] = _sha(synth.candidates(0, 0, 64))
    img = synth.image(0, "V")
    # cfg1: 8x15 RGB, no dither
    o = ob.OracleImage(img, 8, 15)
    o.initialize_tiles()
    out["cfg1_tile_palettes_sha256"] = _sha(o.tile_palettes)
    out["cfg1_seed_palette"] = o.palette[::15].tolist()
    o.recalculate_palettes()
    out["cfg1_palette_sha256"] = _sha(o.palette)
    out["cfg1_map_sha256"] = _sha(o.palette_map)
    out["cfg1_error"] = o.error()
    o.optimize_palette_entry_random(0, 0, synth.candidates(0, 0, 8))
    out["cfg1_palette_after_step_sha256"] = _sha(o.palette)
    out["cfg1_error_after_step"] = o.error()
    # cfg3: dither
    d = ob.OracleImage(img, 8, 15, dither=True)
    d.palette = o.palette
    d.tile_palettes = o.tile_palettes
    d.optimize()
    out["cfg3_map_sha256"] = _sha(d.palette_map)
    out["cfg3_error"] = d.error()
    # cfg2: CIELAB 4x7
    p = ob.OracleImage(img, 4, 7, perceptual_palettes=True)
    p.initialize_tiles()
    p.recalculate_palettes()
    out["cfg2_palette"] = p.palette.tolist()
    out["cfg2_error"] = p.error()
    # cfg4: NES 4x3 dithered
    n = ob.OracleImage(img, 4, 3, dither=True, nes=True)
    n.initialize_tiles()
    n.recalculate_palettes()
    out["cfg4_palette"] = n.palette.tolist()
    out["cfg4_error"] = n.error()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json")
    with open(path, "w") as f:
        json.dump(compute(), f, indent=1, sort_keys=True)
    print("wrote", path)
