"""Writes the inputs tests/golden/reference/gen_reference_vectors.rs reads (raw files, no Rust-side generators to get
wrong): colour pairs, Lab values, a k-means point set, and (source, rendered) image pairs from the oracle's own optimiser
states on the synthetic images the parity tests use.

usage: python tests/golden/write_reference_inputs.py [outdir]      (default tests/golden/reference/inputs)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob          # noqa: E402
from snesimage_b200 import synth          # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    rng = np.random.RandomState(20261018)
    pairs = rng.randint(0, 256, (400, 6)).astype(np.uint8)
    pairs[:20, 3:] = pairs[:20, :3]                      # identical colours: distance 0
    pairs[20:60, 3:] = np.clip(pairs[20:60, :3].astype(int) + rng.randint(-2, 3, (40, 3)), 0, 255)   # near misses
    pairs.tofile(os.path.join(out, "pairs.u8"))
    labs = np.stack([rng.uniform(0, 100, 300), rng.uniform(-90, 90, 300), rng.uniform(-90, 90, 300)], axis=1)
    labs[:8] = [[0, 0, 0], [100, 0, 0], [50, 0, 0], [53.2408, 80.0925, 67.2032], [87.7347, -86.1827, 83.1793], [32.2970, 79.1875, -107.8602],
                [120, 10, 10], [-5, 0, 0]]               # black, white, grey, the sRGB primaries, out-of-gamut values (clamped)
    labs.astype("<f8").tofile(os.path.join(out, "labs.f64"))
    # k-means: the opaque pixels of subpalette 0 of a synthetic image in the reference's gather order would need the
    # reference's tile assignment; a plain seeded point cloud with duplicates exercises the same Lloyd iteration
    pts = np.concatenate([rng.normal(c, 12.0, (400, 3)) for c in ([40, 60, 200], [200, 180, 30], [120, 120, 120], [250, 20, 20])])
    pts = np.clip(np.round(pts), 0, 255)
    pts[5] = pts[0]                                      # two of the first k points identical: cogset's first-k seeding
    pts.astype("<f8").tofile(os.path.join(out, "points.f64"))
    open(os.path.join(out, "k.txt"), "w").write("7\n")
    lines = []
    for seed, family, C, S, dither in [(0, "V", 8, 15, False), (0, "V", 8, 15, True), (3, "T", 4, 7, False), (5, "G", 4, 3, True), (7, "B", 1, 7, False)]:
        rgba = synth.image(seed, family)
        o = ob.OracleImage(rgba, C, S, dither)
        o.initialize_tiles()
        o.recalculate_palettes()
        name = f"{family}{seed}_{C}x{S}{'_d' if dither else ''}"
        rgba.tofile(os.path.join(out, f"{name}_src.rgba"))
        o.as_rgba().tofile(os.path.join(out, f"{name}_dst.rgba"))
        lines.append(f"{name}_src.rgba {name}_dst.rgba")
    lines.append(f"{name}_src.rgba {name}_src.rgba")     # identical images: score 100, error 0
    open(os.path.join(out, "image_pairs.txt"), "w").write("\n".join(lines) + "\n")
    print(f"wrote {len(os.listdir(out))} files to {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "reference", "inputs"))
