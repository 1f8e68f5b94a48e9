"""Pins the oracle against vectors computed by the REAL crates (ssimulacra2 0.5.1 / yuvxyb 0.4.2, palette 0.7.6, cogset
0.2.0) when tests/golden/reference_vectors.json exists -- the output of tests/golden/reference/gen_reference_vectors.rs,
which needs a Rust toolchain this repository's image does not have.  Until someone with cargo runs that recipe the file is
absent and the first test below says PARITY UNPINNED, loudly, instead of passing.

The machinery itself is tested either way: `oracle_document` builds the same document from the oracle, and the comparison
code must accept it (and reject a perturbed copy), so a maintainer who drops the real file in gets a meaningful verdict.
"""
import json
import os
import struct
import warnings

import numpy as np
import pytest

from oracle import binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = os.path.join(HERE, "golden", "reference_vectors.json")
SCORE_TOL = 1e-4      # north star: |delta score| <= 1e-4
LAB_TOL = 2e-5        # f32 Lab components / CIEDE2000 (different cbrt / atan2 implementations differ by ulps)


def _f32(v):
    return {"bits": int(np.float32(v).view(np.uint32)), "value": float(v)}


def _f64(v):
    return {"bits": str(struct.unpack("<Q", struct.pack("<d", float(v)))[0]), "value": float(v)}


def _inputs(tmp):
    import importlib.util
    spec = importlib.util.spec_from_file_location("write_reference_inputs", os.path.join(HERE, "golden", "write_reference_inputs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.main(str(tmp))
    return str(tmp)


def oracle_document(inputs: str) -> dict:
    """What gen_reference_vectors.rs prints, computed by the oracle instead of the crates (same keys, same order)."""
    yux, pal = ob.transfer_luts()
    grid = []
    for r in range(0, 256, 51):
        for g in range(0, 256, 51):
            for b in range(0, 256, 51):
                lab = ob.srgb8_to_lab(r, g, b)
                grid.append({"rgb": [r, g, b], "lab": [_f32(x) for x in lab]})
    pairs = np.fromfile(os.path.join(inputs, "pairs.u8"), np.uint8).reshape(-1, 6)
    ciede = [{"a": p[:3].tolist(), "b": p[3:].tolist(), "d": _f64(ob.cielab(p[:3], p[3:]))} for p in pairs]
    labs = np.fromfile(os.path.join(inputs, "labs.f64"), "<f8").reshape(-1, 3)
    l2s = [{"lab": [_f64(x) for x in v], "rgb": ob.lab_to_srgb8(v).tolist()} for v in labs]
    pts = np.fromfile(os.path.join(inputs, "points.f64"), "<f8").reshape(-1, 3)
    k = int(open(os.path.join(inputs, "k.txt")).read())
    _, centres, assign = ob.kmeans(pts, k)
    clusters = [{"centre": [_f64(x) for x in centres[q]], "members": np.flatnonzero(assign == q).tolist()} for q in range(k)]
    scores = []
    for line in open(os.path.join(inputs, "image_pairs.txt")).read().split("\n"):
        if not line.strip():
            continue
        s, d = line.split()
        src = np.fromfile(os.path.join(inputs, s), np.uint8).reshape(256, 256, 4)
        dst = np.fromfile(os.path.join(inputs, d), np.uint8).reshape(256, 256, 4)
        sc = ob.ssimulacra2(src, dst)
        scores.append({"src": s, "dst": d, "ssimulacra2": _f64(sc), "error": _f64(100.0 - sc)})
    return {"crates": {"oracle": "restatement"}, "eotf_yuvxyb": [_f32(v) for v in yux], "eotf_palette": [_f32(v) for v in pal],
            "lab_grid": grid, "ciede2000": ciede, "lab_to_srgb8": l2s, "kmeans": {"k": k, "clusters": clusters}, "ssimulacra2": scores}


def compare(ref: dict, inputs: str) -> dict:
    """Oracle against a reference document.  Returns a report; raises AssertionError where a stated bound is broken.
    If the reference's transfer tables differ from the built-in ones they are injected into the oracle first (that is
    what they are injectable for) and the report says how far apart they were."""
    rep = {}
    yux = np.array([e["bits"] for e in ref["eotf_yuvxyb"]], np.uint32).view(np.float32)
    pal = np.array([e["bits"] for e in ref["eotf_palette"]], np.uint32).view(np.float32)
    try:
        oy, op = ob.transfer_luts()
        rep["eotf_yuvxyb_bit_identical"] = bool(np.array_equal(oy.view(np.uint32), yux.view(np.uint32)))
        rep["eotf_palette_bit_identical"] = bool(np.array_equal(op.view(np.uint32), pal.view(np.uint32)))
        rep["eotf_yuvxyb_max_abs_diff"] = float(np.max(np.abs(oy - yux)))
        rep["eotf_palette_max_abs_diff"] = float(np.max(np.abs(op - pal)))
        assert rep["eotf_yuvxyb_max_abs_diff"] <= 1e-4 and rep["eotf_palette_max_abs_diff"] <= 1e-6, rep   # same function at all?
        ob.set_transfer_luts(yux, pal)
        mine = oracle_document(inputs)
    finally:
        ob.set_transfer_luts(None, None)
    lab_r = np.array([[c["bits"] for c in e["lab"]] for e in ref["lab_grid"]], np.uint32).view(np.float32)
    lab_m = np.array([[c["bits"] for c in e["lab"]] for e in mine["lab_grid"]], np.uint32).view(np.float32)
    rep["lab_max_abs_diff"] = float(np.max(np.abs(lab_r - lab_m)))
    rep["lab_bit_identical_fraction"] = float(np.mean(lab_r.view(np.uint32) == lab_m.view(np.uint32)))
    assert rep["lab_max_abs_diff"] <= LAB_TOL * 100, rep          # Lab components reach 100
    d_r = np.array([e["d"]["value"] for e in ref["ciede2000"]])
    d_m = np.array([e["d"]["value"] for e in mine["ciede2000"]])
    rep["ciede2000_max_abs_diff"] = float(np.max(np.abs(d_r - d_m)))
    assert rep["ciede2000_max_abs_diff"] <= LAB_TOL * 10, rep
    bad = [i for i, (a, b) in enumerate(zip(ref["lab_to_srgb8"], mine["lab_to_srgb8"])) if a["rgb"] != b["rgb"]]
    rep["lab_to_srgb8_mismatches"] = len(bad)
    assert not bad, (bad[:5], rep)
    assert ref["kmeans"]["k"] == mine["kmeans"]["k"]
    for q, (a, b) in enumerate(zip(ref["kmeans"]["clusters"], mine["kmeans"]["clusters"])):
        assert a["members"] == b["members"], f"cluster {q}: member lists differ"
        ca, cb = [c["value"] for c in a["centre"]], [c["value"] for c in b["centre"]]
        assert np.allclose(ca, cb, rtol=0, atol=1e-9, equal_nan=True), (q, ca, cb)
    rep["kmeans_clusters"] = len(ref["kmeans"]["clusters"])
    e_r = np.array([e["error"]["value"] for e in ref["ssimulacra2"]])
    e_m = np.array([e["error"]["value"] for e in mine["ssimulacra2"]])
    rep["ssimulacra2_max_abs_diff"] = float(np.max(np.abs(e_r - e_m)))
    assert rep["ssimulacra2_max_abs_diff"] <= SCORE_TOL, rep
    return rep


def test_oracle_against_reference_vectors(tmp_path):
    if not os.path.exists(VECTORS):
        msg = ("PARITY UNPINNED: tests/golden/reference_vectors.json is absent -- no Rust toolchain in this image to run "
               "tests/golden/reference/gen_reference_vectors.rs against ssimulacra2 0.5.1 / palette 0.7.6 / cogset 0.2.0. "
               "The oracle is pinned only to published known answers (tests/test_oracle.py) and to lib.rs itself.")
        warnings.warn(msg)
        pytest.skip(msg)
    with open(VECTORS) as f:
        ref = json.load(f)
    rep = compare(ref, _inputs(tmp_path))
    print("oracle vs reference crates:", json.dumps(rep, indent=1))


def test_comparison_machinery_accepts_the_oracle_and_rejects_a_perturbed_copy(tmp_path):
    inputs = _inputs(tmp_path)
    doc = json.loads(json.dumps(oracle_document(inputs)))          # through JSON, as the real file would come
    rep = compare(doc, inputs)
    assert rep["eotf_yuvxyb_bit_identical"] and rep["eotf_palette_bit_identical"] and rep["lab_bit_identical_fraction"] == 1.0
    assert rep["ssimulacra2_max_abs_diff"] == 0.0 and rep["ciede2000_max_abs_diff"] == 0.0
    assert doc["ssimulacra2"][-1]["error"]["value"] == 0.0         # identical images: error() == 0 exactly
    bad = json.loads(json.dumps(doc))
    bad["ssimulacra2"][0]["error"]["value"] += 5e-4                # five times the stated tolerance
    with pytest.raises(AssertionError):
        compare(bad, inputs)
    bad = json.loads(json.dumps(doc))
    bad["kmeans"]["clusters"][2]["members"][0] += 1
    with pytest.raises(AssertionError):
        compare(bad, inputs)
    # a transfer table that differs in its last bits is injected, and the scores follow it (that is what injection is for)
    bad = json.loads(json.dumps(doc))
    for e in bad["eotf_yuvxyb"][100:140]:
        e["bits"] += 3
    with pytest.raises(AssertionError):
        compare(bad, inputs)                                       # ... so the untouched scores of `bad` no longer match


@pytest.mark.gpu
def test_transfer_lut_injection_reaches_the_gpu_path(ctx):
    """The same perturbed tables injected into the oracle and into the library: source planes still bit-identical, Lab
    planes within tolerance, error() equal -- a verified table swaps in without touching a kernel."""
    from snesimage_b200 import engine, synth
    yux, pal = ob.transfer_luts()
    yux2 = (yux.view(np.uint32) + np.where(np.arange(256) > 10, 5, 0).astype(np.uint32)).view(np.float32)
    pal2 = (pal.view(np.uint32) + np.where(np.arange(256) > 10, 3, 0).astype(np.uint32)).view(np.float32)
    rgba = synth.image(12, "V")
    try:
        ob.set_transfer_luts(yux2, pal2)
        ctx.set_transfer_luts(yux2, pal2)
        cfg = engine.Config(subpalette_count=4, subpalette_size=7, perceptual_palettes=True)
        g = engine.OptimizedImage(ctx, rgba, cfg)
        o = ob.OracleImage(rgba, 4, 7, False, True, False)
        xyb, mu1, s11 = g.debug_planes()
        assert np.array_equal(xyb.view(np.uint32), ob.xyb_pyramid(rgba).view(np.uint32))
        omu1, _ = ob.source_planes(rgba)
        assert np.array_equal(mu1.view(np.uint32), omu1.view(np.uint32))
        lab = g.debug_lab().reshape(256, 256, 3)
        want = np.stack([ob.srgb8_to_lab(*rgba[y, x, :3]) for y, x in [(0, 0), (17, 200), (255, 255), (128, 3)]])
        assert np.max(np.abs(lab[[0, 17, 255, 128], [0, 200, 255, 3]] - want)) <= 1e-4
        for im in (g, o):
            im.initialize_tiles()
        o.tile_palettes = g.tile_palettes
        o.palette = g.palette
        o.palette_map = g.palette_map
        assert abs(g.error() - o.error()) <= 1e-8
        g.close()
        # and the planes did move with the tables
        ob.set_transfer_luts(None, None)
        assert not np.array_equal(xyb.view(np.uint32), ob.xyb_pyramid(rgba).view(np.uint32))
    finally:
        ob.set_transfer_luts(None, None)
        ctx.set_transfer_luts(None, None)
