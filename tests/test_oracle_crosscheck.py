"""CPU tests (no GPU): the oracle's restated third-party arithmetic against INDEPENDENT implementations that happen to
be in this image (OpenCV, SciPy).  None of them is the reference's crate, so this does not lift "parity unpinned"
(DESIGN.md section 2); what it rules out is a mis-recalled white point, matrix, sigma or Lloyd update in
oracle/constants_unverified.h -- errors far larger than the tolerances below.

  palette 0.7.6 sRGB -> Lab<D65, f32>   (lib.rs:101-103, 344-346, 1092-1099)   vs  cv2.cvtColor(..., COLOR_RGB2Lab)
  ssimulacra2 recursive Gaussian, sigma 1.5 (error(), lib.rs:503-548)          vs  scipy.ndimage.gaussian_filter
  cogset 0.2.0 Kmeans, first-k centres   (lib.rs:119-133, 348-368)             vs  scipy.cluster.vq.kmeans2(minit="matrix")
"""
import numpy as np
import pytest

from oracle import binding as ob

cv2 = pytest.importorskip("cv2")
ndi = pytest.importorskip("scipy.ndimage")
vq = pytest.importorskip("scipy.cluster.vq")


def test_lab_against_opencv():
    rng = np.random.RandomState(1)
    cols = rng.randint(0, 256, (4000, 3)).astype(np.uint8)
    cols[:4] = [(0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 0, 255)]
    ours = np.array([ob.srgb8_to_lab(*c) for c in cols])
    theirs = cv2.cvtColor((cols.astype(np.float32) / 255.0)[None], cv2.COLOR_RGB2Lab)[0]
    # OpenCV's float path uses spline tables for the transfer curve and the cube root (a few 0.1 units); a wrong white
    # point or primaries matrix would move a, b by several units
    assert np.abs(ours - theirs).max() < 0.4
    assert np.abs(ours[:2] - [(0, 0, 0), (100, 0, 0)]).max() < 0.02
    back = np.array([ob.lab_to_srgb8(l) for l in ours])
    assert np.array_equal(back, cols)


def test_lab_to_srgb8_against_opencv():
    rng = np.random.RandomState(2)
    lab = np.stack([rng.uniform(5, 95, 2000), rng.uniform(-40, 40, 2000), rng.uniform(-40, 40, 2000)], 1).astype(np.float32)
    theirs = cv2.cvtColor(lab[None], cv2.COLOR_Lab2RGB)[0]
    inside = np.all((theirs > 0.02) & (theirs < 0.98), axis=1)     # away from gamut clipping, where conventions differ
    ours = np.array([ob.lab_to_srgb8(l) for l in lab[inside]]).astype(int)
    assert inside.sum() > 500
    assert np.abs(ours - np.round(theirs[inside] * 255.0)).max() <= 2


def test_recursive_gaussian_against_scipy():
    rng = np.random.RandomState(3)
    plane = rng.rand(96, 96).astype(np.float32)
    ours = ob.blur_plane(plane)
    theirs = ndi.gaussian_filter(plane.astype(np.float64), 1.5, mode="constant", truncate=6.0)
    # the 3-section recursive filter approximates the sigma = 1.5 Gaussian to ~2e-3 on [0, 1] data, zero-padded like
    # scipy's mode="constant"; sigma = 1.4 or 1.6 would be off by > 1e-2
    assert np.abs(ours - theirs).max() < 3e-3
    for wrong in (1.4, 1.6):
        assert np.abs(ours - ndi.gaussian_filter(plane.astype(np.float64), wrong, mode="constant", truncate=6.0)).max() > 6e-3


@pytest.mark.parametrize("k,n,seed", [(5, 500, 1), (8, 1024, 2), (15, 4000, 3)])
def test_lloyd_against_scipy(k, n, seed):
    rng = np.random.RandomState(seed)
    pts = rng.rand(n, 3) * 255.0
    iters, centres, labels = ob.kmeans(pts, k)
    cen, lab = vq.kmeans2(pts, pts[:k].copy(), iter=max(int(iters), 1) + 50, minit="matrix")
    assert np.abs(np.asarray(centres) - cen).max() < 1e-9
    assert np.array_equal(np.asarray(labels), lab)


def test_srgb_eotf_against_iec_formula():
    """yuvxyb linearises with the H.273 constants (alpha 1.0550107..., beta 0.0030412825...), which differ from the rounded
    IEC 61966-2-1 ones (1.055, 0.04045) in the sixth digit: all 256 levels agree to 1e-5, and a gamma-2.2 curve would not."""
    v = np.arange(256) / 255.0
    iec = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    ours = np.array([ob.lib().ora_srgb_eotf(float(np.float32(x))) for x in v])
    assert np.abs(ours - iec).max() < 1e-5
    assert np.abs(ours - v ** 2.2).max() > 5e-3
    assert np.all(np.diff(ours) > 0)
