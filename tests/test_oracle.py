"""CPU tests (no GPU): pin the oracle against published known answers and against what
/root/reference/src/lib.rs alone determines.  The reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), and its crates are not vendored, so third-party arithmetic is pinned to the
published standards only: CIEDE2000 to the 34 Sharma/Wu/Dalal pairs, Lab to D65 identities, the
recursive Gaussian to its defining properties, SSIMULACRA2 to its fixed points.  tests/golden/ holds
vectors generated from the oracle itself (regression pins, with the generating script)."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from snesimage_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))

# Sharma, Wu, Dalal (2005), Table 1: L1 a1 b1 L2 a2 b2 dE00
SHARMA = [
    (50.0000, 2.6772, -79.7751, 50.0000, 0.0000, -82.7485, 2.0425),
    (50.0000, 3.1571, -77.2803, 50.0000, 0.0000, -82.7485, 2.8615),
    (50.0000, 2.8361, -74.0200, 50.0000, 0.0000, -82.7485, 3.4412),
    (50.0000, -1.3802, -84.2814, 50.0000, 0.0000, -82.7485, 1.0000),
    (50.0000, -1.1848, -84.8006, 50.0000, 0.0000, -82.7485, 1.0000),
    (50.0000, -0.9009, -85.5211, 50.0000, 0.0000, -82.7485, 1.0000),
    (50.0000, 0.0000, 0.0000, 50.0000, -1.0000, 2.0000, 2.3669),
    (50.0000, -1.0000, 2.0000, 50.0000, 0.0000, 0.0000, 2.3669),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0009, 7.1792),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0010, 7.1792),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0011, 7.2195),
    (50.0000, 2.4900, -0.0010, 50.0000, -2.4900, 0.0012, 7.2195),
    (50.0000, -0.0010, 2.4900, 50.0000, 0.0009, -2.4900, 4.8045),
    (50.0000, -0.0010, 2.4900, 50.0000, 0.0010, -2.4900, 4.8045),
    (50.0000, -0.0010, 2.4900, 50.0000, 0.0011, -2.4900, 4.7461),
    (50.0000, 2.5000, 0.0000, 50.0000, 0.0000, -2.5000, 4.3065),
    (50.0000, 2.5000, 0.0000, 73.0000, 25.0000, -18.0000, 27.1492),
    (50.0000, 2.5000, 0.0000, 61.0000, -5.0000, 29.0000, 22.8977),
    (50.0000, 2.5000, 0.0000, 56.0000, -27.0000, -3.0000, 31.9030),
    (50.0000, 2.5000, 0.0000, 58.0000, 24.0000, 15.0000, 19.4535),
    (50.0000, 2.5000, 0.0000, 50.0000, 3.1736, 0.5854, 1.0000),
    (50.0000, 2.5000, 0.0000, 50.0000, 3.2972, 0.0000, 1.0000),
    (50.0000, 2.5000, 0.0000, 50.0000, 1.8634, 0.5757, 1.0000),
    (50.0000, 2.5000, 0.0000, 50.0000, 3.2592, 0.3350, 1.0000),
    (60.2574, -34.0099, 36.2677, 60.4626, -34.1751, 39.4387, 1.2644),
    (63.0109, -31.0961, -5.8663, 62.8187, -29.7946, -4.0864, 1.2630),
    (61.2901, 3.7196, -5.3901, 61.4292, 2.2480, -4.9620, 1.8731),
    (35.0831, -44.1164, 3.7933, 35.0232, -40.0716, 1.5901, 1.8645),
    (22.7233, 20.0904, -46.6940, 23.0331, 14.9730, -42.5619, 2.0373),
    (36.4612, 47.8580, 18.3852, 36.2715, 50.5065, 21.2231, 1.4146),
    (90.8027, -2.0831, 1.4410, 91.1528, -1.6435, 0.0447, 1.4441),
    (90.9257, -0.5406, -0.9208, 88.6381, -0.8985, -0.7239, 1.5381),
    (6.7747, -0.2908, -2.4247, 5.8714, -0.0985, -2.2286, 0.6377),
    (2.0776, 0.0795, -1.1350, 0.9033, -0.0636, -0.5514, 0.9082),
]


@pytest.mark.parametrize("row", SHARMA)
def test_ciede2000_sharma_pairs(row):
    l1, l2, want = row[0:3], row[3:6], row[6]
    assert abs(ob.ciede2000_f64(l1, l2) - want) < 5e-5
    assert abs(ob.ciede2000_f64(l2, l1) - want) < 5e-5
    assert abs(ob.ciede2000_f32(l1, l2) - want) < 2e-3  # the f32 form the reference uses


def test_lab_d65_identities():
    assert np.allclose(ob.srgb8_to_lab(255, 255, 255), [100.0, 0.0, 0.0], atol=2e-2)
    assert np.allclose(ob.srgb8_to_lab(0, 0, 0), [0.0, 0.0, 0.0], atol=1e-6)
    # published sRGB primaries in Lab (D65): red (53.24, 80.09, 67.20), green (87.73, -86.18, 83.18), blue (32.30, 79.19, -107.86)
    assert np.allclose(ob.srgb8_to_lab(255, 0, 0), [53.24, 80.09, 67.20], atol=0.02)
    assert np.allclose(ob.srgb8_to_lab(0, 255, 0), [87.73, -86.18, 83.18], atol=0.02)
    assert np.allclose(ob.srgb8_to_lab(0, 0, 255), [32.30, 79.19, -107.86], atol=0.02)
    # greys have no chroma
    for v in (1, 17, 128, 200):
        lab = ob.srgb8_to_lab(v, v, v)
        assert abs(lab[1]) < 2e-2 and abs(lab[2]) < 2e-2


def test_lab_round_trip_to_srgb8():
    rng = np.random.RandomState(0)
    for r, g, b in rng.randint(0, 256, (300, 3)):
        lab = ob.srgb8_to_lab(r, g, b).astype(np.float64)
        assert np.array_equal(ob.lab_to_srgb8(lab), [r, g, b])
    assert np.array_equal(ob.lab_to_srgb8([float("nan")] * 3), [0, 0, 0])


def test_snes_color_primitives():
    # lib.rs:662-669, u8 arithmetic; lib.rs:679-681
    for c in range(32):
        assert ob.snes_as_rgba([c, c, c])[0] == c * 8 + c // 4
    assert list(ob.snes_as_rgba([31, 0, 16])) == [255, 0, 132, 255]
    assert list(ob.snes_as_rgba([32, 32, 32])) == [8, 8, 8, 255]  # release-mode wrap of 32*8
    assert ob.lib().ora_snes_as_u16(np.array([1, 2, 3], np.uint8)) == 1 + (2 << 5) + (3 << 10)
    # NES table: 56 entries, 54 distinct, out-of-range index -> black (lib.rs:743)
    table = np.stack([ob.nes_color(i) for i in range(56)])
    assert len({tuple(t) for t in table}) == 54
    assert list(table[0]) == [13, 13, 13] and list(table[55]) == [23, 24, 23]
    assert list(ob.nes_color(56)) == [0, 0, 0]
    assert list(ob.new_nes_only([31, 31, 31])) == [31, 31, 31]
    assert list(ob.new_nes_only([0, 0, 0])) == [0, 0, 0]


def test_red_mean_formula_and_integer_key():
    # lib.rs:1080-1088 verbatim in numpy f64, and the int32 key the GPU kernels use (= 512 d^2)
    rng = np.random.RandomState(1)
    a = rng.randint(0, 256, (5000, 3))
    b = rng.randint(0, 256, (5000, 3))
    rm = (a[:, 0] + b[:, 0]) / 2.0
    d = (a - b).astype(np.float64)
    want = np.sqrt(((512.0 + rm) * d[:, 0] ** 2) / 256.0 + 4.0 * d[:, 1] ** 2 + ((767.0 - rm) * d[:, 2] ** 2) / 256.0)
    got = np.array([ob.red_mean(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, want)
    rs = a[:, 0] + b[:, 0]
    di = a - b
    key = (1024 + rs) * di[:, 0] ** 2 + 2048 * di[:, 1] ** 2 + (1534 - rs) * di[:, 2] ** 2
    assert np.array_equal(np.sqrt(key / 512.0), want)
    order_f = np.argsort(want, kind="stable")
    order_k = np.argsort(key, kind="stable")
    assert np.array_equal(order_f, order_k)
    assert key.max() < 2 ** 31


def test_expanded_red_mean_key_and_rounding_rule_of_the_dither_kernel():
    # The expansion behind the dither kernel's packed key (snesimage_b200/csrc/dither_core.h; the packed form itself, with its
    # per-pixel shift and the entry number in the low bits, is checked in tests/test_dither_core.py): the integer key with the
    # target-only terms dropped,
    #   key'' = C0 + A r - R (r^2 + b^2) - 4096 G g + C1 b + 2B (r b)   (int32, wrap-around arithmetic)
    # ranks the entries of a pixel exactly like the full key (first strict minimum).  Exhaustive over the
    # corners and a random sample of (target, entry) pairs, in int32 with overflow wrapping like the GPU.
    rng = np.random.RandomState(7)
    corners = np.array([[r, g, b] for r in (0, 1, 127, 128, 254, 255) for g in (0, 128, 255) for b in (0, 1, 128, 255)])
    tg = np.concatenate([corners, rng.randint(0, 256, (3000, 3))]).astype(np.int64)
    en = np.concatenate([corners, rng.randint(0, 256, (3000, 3))]).astype(np.int64)
    r, g, b = tg[:, None, 0], tg[:, None, 1], tg[:, None, 2]
    R, G, B = en[None, :, 0], en[None, :, 1], en[None, :, 2]
    key = (1024 + r + R) * (R - r) ** 2 + 2048 * (G - g) ** 2 + (1534 - r - R) * (B - b) ** 2
    const = r ** 3 + 1024 * r ** 2 + 2048 * g ** 2 + 1534 * b ** 2 - r * b ** 2
    C0 = (1024 + R) * R * R + 2048 * G * G + (1534 - R) * B * B
    A = -R * R - 2048 * R - B * B
    C1 = -2 * B * (1534 - R)
    with np.errstate(over="ignore"):
        i32 = lambda v: v.astype(np.int32)  # noqa: E731  (wraps modulo 2^32 like the device arithmetic)
        k2 = i32(C0) + i32(A) * i32(r) + i32(-R) * i32(r * r + b * b) + i32(-4096 * G) * i32(g) + i32(C1) * i32(b) + i32(2 * B) * i32(r * b)
    assert np.array_equal(k2.astype(np.int64), key - const)          # exact, no overflow in the final value
    assert np.abs(key - const).max() < 2 ** 31 and np.abs(C0).max() < 2 ** 31
    # ranking of 15-entry subpalettes: first strict minimum of key'' == first strict minimum of key
    for trial in range(200):
        sub = rng.choice(len(en), 15, replace=trial % 2 == 0)  # with replacement: duplicate entries -> ties
        assert np.array_equal(np.argmin(key[:, sub], axis=1), np.argmin(k2[:, sub], axis=1))

    # clamp(0, 255).round() (half away from zero, lib.rs:773-778) as truncate + exact-fraction test + integer clamp
    eps = np.finfo(np.float64).eps
    t = np.concatenate([rng.uniform(-300, 600, 20000), np.arange(-3, 259) + 0.5, np.arange(-3, 259) + 0.5 - 64 * eps,
                        np.arange(-3, 259) - 0.5 + 64 * eps, [0.49999999999999994, 254.99999999999997, -0.0, 1e300, -1e300]])
    c = np.clip(t, 0.0, 255.0)
    want = np.where(c - np.floor(c) >= 0.5, np.floor(c) + 1, np.floor(c)).astype(np.int64)  # exact half-away rule for c >= 0
    tz = np.trunc(np.clip(t, -2.0 ** 31, 2.0 ** 31 - 1))             # cvt.rzi.s32.f64 saturates
    got = np.clip(tz + (t - tz >= 0.5), 0, 255).astype(np.int64)
    assert np.array_equal(got, want)
    for v in (0.5, 1.5, 2.5, 254.5, 0.49999999999999994, -0.5, 255.5):  # against the oracle's own rounding
        col = np.array([[0, 0, 0], [0, 0, 1]], np.uint8)               # as_rgba blue 0 and 8
        rounded = int(got[np.where(t == v)[0][0]]) if (t == v).any() else None
        if rounded is not None:
            assert ob.closest_color_index(col, [0.0, 0.0, v]) == (1 if abs(rounded - 8) < abs(rounded - 0) else 0)


def test_closest_color_strict_first_min_and_rounding():
    colors = np.array([[0, 0, 0], [10, 10, 10], [10, 10, 10], [31, 31, 31]], np.uint8)
    assert ob.closest_color_index(colors, [82.0, 82.0, 82.0]) == 1          # tie between 1 and 2 -> lowest index
    assert ob.closest_color_index(colors, [-50.0, -1.0, 0.0]) == 0          # clamp
    assert ob.closest_color_index(colors, [300.0, 999.0, 255.4]) == 3
    two = np.array([[0, 0, 0], [0, 0, 1]], np.uint8)                        # rgba 0 and 8 in blue
    assert ob.closest_color_index(two, [0.0, 0.0, 3.5]) == 0                # 3.5 rounds to 4: tie -> index 0
    assert ob.closest_color_index(two, [0.0, 0.0, 4.5]) == 1                # 4.5 rounds to 5 (half away from zero)


def test_optimize_without_dither_is_per_pixel_nearest():
    rgba = synth.image(5, "T")
    o = ob.OracleImage(rgba, 4, 7)
    o.palette = synth.random_palette(5, 4, 7)
    o.tile_palettes = synth.random_tile_palettes(5, 4)
    o.optimize()
    pm, pal, tp = o.palette_map, o.palette, o.tile_palettes
    rng = np.random.RandomState(2)
    for y, x in rng.randint(0, 256, (300, 2)):
        sub = int(tp[(y // 8) * 32 + x // 8]) * 7
        want = ob.closest_color_index(pal[sub:sub + 7], rgba[y, x, :3].astype(np.float64))
        assert pm[y, x] == (want if rgba[y, x, 3] > 0 else 0)
    assert (pm[rgba[..., 3] == 0] == 0).all()
    out = o.as_rgba()
    assert (out[rgba[..., 3] == 0] == 0).all()          # transparent -> (0,0,0,0), lib.rs:551-558
    assert (out[rgba[..., 3] > 0][:, 3] == 255).all()


def test_dither_first_row_matches_python_loop():
    """Independent restatement of lib.rs:425-501 for row 0 (only the E term is live there)."""
    rgba = synth.image(6, "V")
    o = ob.OracleImage(rgba, 1, 5, dither=True)
    pal = synth.random_palette(6, 1, 5)
    o.palette = pal
    o.optimize()
    err = np.zeros(3)
    for x in range(256):
        target = rgba[0, x, :3].astype(np.float64) + err
        k = ob.closest_color_index(pal, target)
        assert o.palette_map[0, x] == k
        err = (target - ob.snes_as_rgba(pal[k])[:3].astype(np.float64)) * 0.8 * (7.0 / 16.0)


def test_kmeans_restatement():
    # two well-separated blobs, first-k initial centres both in blob A
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [100, 100, 100], [101, 100, 100], [100, 101, 100]], np.float64)
    it, centres, assign = ob.kmeans(pts, 2)
    assert it >= 1
    assert list(assign) == [0, 0, 0, 1, 1, 1] or list(assign) == [1, 1, 1, 0, 0, 0]
    assert np.allclose(sorted(centres[:, 0]), [1 / 3, 100 + 1 / 3])
    # cogset asserts 2 <= k < n
    assert ob.kmeans(pts, 1)[0] == -1 and ob.kmeans(pts, 6)[0] == -1
    # identical initial points -> one empty cluster -> NaN centre (0 * inf)
    pts2 = np.array([[5, 5, 5], [5, 5, 5], [9, 9, 9], [1, 1, 1]], np.float64)
    _, c2, a2 = ob.kmeans(pts2, 2)
    assert np.isnan(c2[1]).all() and (a2 == 0).all()


def test_gaussian_blur_properties():
    n2, d1, radius = ob.gaussian_coeffs()
    assert radius == 5
    # impulse response: symmetric, sums to ~1, peak at the centre, close to a sigma=1.5 Gaussian
    img = np.zeros((64, 64), np.float32)
    img[32, 32] = 1.0
    out = ob.blur_plane(img)
    assert abs(out.sum() - 1.0) < 2e-3
    assert np.unravel_index(np.argmax(out), out.shape) == (32, 32)
    assert np.allclose(out[32, 33:40], out[32, 31:24:-1], atol=1e-6)
    assert np.allclose(out, out.T, atol=1e-6)
    x = np.arange(-16, 17)
    g = np.exp(-x ** 2 / (2 * 1.5 ** 2))
    g /= g.sum()
    row = out[32, 16:49] / out[32].sum()
    assert np.max(np.abs(row - g)) < 5e-3
    # constants stay constant away from the zero-padded border
    flat = ob.blur_plane(np.full((64, 64), 0.7, np.float32))
    assert np.allclose(flat[20:44, 20:44], 0.7, atol=2e-3)


def test_xyb_and_transfer_known_values():
    assert ob.lib().ora_srgb_eotf(0.0) == 0.0
    assert abs(ob.lib().ora_srgb_eotf(1.0) - 1.0) < 1e-6
    assert abs(ob.lib().ora_srgb_eotf(0.5) - 0.21404114) < 1e-5   # IEC 61966-2-1 (yuvxyb uses the H.273 constants)
    for v in (0.0, 1e-6, 0.001, 0.5, 1.0, 27.0, 1e10):
        assert abs(ob.lib().ora_cbrtf(v) - np.cbrt(np.float32(v))) <= 1e-6 * max(1.0, np.cbrt(v))
    xyb = np.zeros(3, np.float32)
    ob.lib().ora_linear_rgb_to_xyb(np.zeros(3, np.float32), xyb)
    assert np.allclose(xyb, 0.0, atol=1e-6)                        # black -> XYB origin
    ob.lib().ora_linear_rgb_to_xyb(np.ones(3, np.float32), xyb)
    assert abs(xyb[0]) < 1e-6 and abs(xyb[1] - xyb[2]) < 1e-6      # white: no X, Y == B


def test_ssimulacra2_fixed_points_and_ordering():
    img = synth.image(9, "V")
    assert ob.ssimulacra2(img, img) == 100.0
    rng = np.random.RandomState(3)
    prev = 100.0
    for amp in (2, 6, 16, 40):     # more distortion -> lower score
        n = img.copy()
        n[..., :3] = np.clip(n[..., :3].astype(int) + rng.randint(-amp, amp + 1, (256, 256, 3)), 0, 255)
        s = ob.ssimulacra2(img, n)
        assert s < prev
        prev = s
    # alpha is ignored on the source side (lib.rs:506-516)
    a = img.copy()
    a[..., 3] = 0
    assert ob.ssimulacra2(a, img) == 100.0
    _, avg = ob.ssimulacra2(img, n, want_avg=True)
    assert avg.shape == (6, 18) and (avg >= 0).all()


def test_as_json_layout():
    rgba = synth.image(7, "T")
    o = ob.OracleImage(rgba, 3, 5)
    o.palette = synth.random_palette(7, 3, 5)
    o.tile_palettes = synth.random_tile_palettes(7, 3)
    o.optimize()
    doc = o.as_json()
    pal = np.array(doc["palette"]).reshape(3, 16)
    assert (pal[:, 0] == 0).all() and (pal[:, 6:] == 0).all()
    p5 = o.palette.astype(int)
    assert np.array_equal(pal[:, 1:6].ravel(), p5[:, 0] + (p5[:, 1] << 5) + (p5[:, 2] << 10))
    tiles = np.array(doc["tiles"])
    assert tiles.shape == (1024, 64)
    t = 5 * 32 + 9
    block = o.palette_map[40:48, 72:80].astype(int) + 1
    block[rgba[40:48, 72:80, 3] == 0] = 0
    assert np.array_equal(tiles[t].reshape(8, 8), block)
    assert doc["tile_palettes"] == o.tile_palettes.tolist()


def test_golden_regression_vectors():
    """tests/golden/oracle_vectors.json is produced by tests/golden/make_golden.py from this oracle; it
    pins the oracle against silent drift (it is NOT a reference-derived vector: the reference has none)."""
    with open(os.path.join(HERE, "golden", "oracle_vectors.json")) as f:
        gold = json.load(f)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    now = mg.compute()
    assert now.keys() == gold.keys()
    for k in gold:
        if isinstance(gold[k], float):
            assert abs(now[k] - gold[k]) <= 1e-9 * max(1.0, abs(gold[k])), k
        else:
            assert now[k] == gold[k], k


def test_ciede2000_lightness_lower_bound_of_the_cielab_candidate_kernel():
    """k_assign_pyr<1> (assign_delta.cuh) decides a pixel without the CIEDE2000 formula when |dL| / 1.75 exceeds the pixel's threshold
    times 1.001 plus 1e-3.  That is safe iff the f32 distance d satisfies |dL| / 1.75 <= 1.001 d + 1e-3 for every pair the kernel can
    meet: candidate = a BGR555 colour through as_rgba, pixel = any sRGB8 colour.  Mathematically dE00 >= |dL| / SL with SL <= 1.747
    (the chroma / hue terms form a positive-semidefinite quadratic because |RT| < 2); here the oracle's f32 evaluation is asked."""
    rng = np.random.RandomState(11)
    n = 60000
    cand5 = rng.randint(0, 32, (n, 3))
    pix = rng.randint(0, 256, (n, 3))
    # the corners of the bound: saturated blues and purples (where RT is largest), greys, near-identical and extreme-lightness pairs
    pix[:4000, 2] = rng.randint(200, 256, 4000)
    pix[:4000, :2] = rng.randint(0, 60, (4000, 2))
    cand5[:2000] = np.stack([rng.randint(0, 8, 2000), rng.randint(0, 8, 2000), rng.randint(24, 32, 2000)], axis=1)
    pix[4000:6000] = np.repeat(rng.randint(0, 256, (2000, 1)), 3, axis=1)
    cand5[5000:7000] = np.repeat(rng.randint(0, 32, (2000, 1)), 3, axis=1)
    pix[7000:7100] = 255
    cand5[7000:7100] = 0
    pix[7100:7200] = 0
    cand5[7100:7200] = 31
    worst = np.inf
    for c5, p in zip(cand5, pix):
        c8 = ob.snes_as_rgba(c5.astype(np.uint8))[:3]
        lc, lp = ob.srgb8_to_lab(*[int(v) for v in c8]), ob.srgb8_to_lab(*[int(v) for v in p])
        d = ob.ciede2000_f32(lc, lp)                       # argument order of lib.rs:783: (palette colour, target)
        lb = abs(float(lc[0]) - float(lp[0])) / 1.75
        assert lb <= 1.001 * d + 1e-3, (c5, p, d, lb)
        if lb > 1.0:
            worst = min(worst, d / lb)
    assert worst >= 1.0     # the bound itself, without the kernel's margin, on every pair with a lightness difference above 1.75
