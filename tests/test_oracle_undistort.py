"""Undistortion: the restatement of OpenCV's fixed-point undistort maps + 8-bit bilinear remap
(oracle/undistort.cpp) must reproduce what OpenCV itself produced -- tests/golden/undistort.npz, recorded from cv2 by
scripts/make_golden_undistort.py -- bit for bit, and live against cv2 where it can be imported."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "scripts"))
GOLDEN = os.path.join(HERE, "golden", "undistort.npz")


def golden():
    z = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {k.split("/")[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names}


G = golden()
SMALL = sorted(n for n in G if "map1" in G[n])
FULL = sorted(n for n in G if "sha256_map1" in G[n])


def texture(seed, w, h):  # same generator as scripts/make_golden_undistort.py (digest cases regenerate their input)
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = 128 + 60 * np.sin(x / 5.0 + seed) * np.cos(y / 7.0) + 40 * rng.standard_normal((h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("name", SMALL)
def test_maps_and_remap_equal_opencv_golden(oracle, name):
    c = G[name]
    h, w = c["map2"].shape
    m1, m2 = oracle.init_undistort_rectify_map(c["K"], c["dist"], c["Kout"], w, h)
    assert np.array_equal(m1, c["map1"]) and np.array_equal(m2, c["map2"])
    assert np.array_equal(oracle.remap_u8(c["image"], c["map1"], c["map2"]), c["undistorted"])
    if name == "d2_small_full":  # alpha = 1: part of the output looks outside the source (border taps are exercised)
        assert (c["map1"][..., 0].min() < 0) or (c["map1"][..., 0].max() >= c["in_wh"][0] - 1)


@pytest.mark.parametrize("name", FULL)
def test_full_size_digests(oracle, name):
    c = G[name]
    ow, oh = (int(v) for v in c["out_wh"])
    iw, ih = (int(v) for v in c["in_wh"])
    m1, m2 = oracle.init_undistort_rectify_map(c["K"], c["dist"], c["Kout"], ow, oh)
    assert np.array_equal(sha(m1), c["sha256_map1"]) and np.array_equal(sha(m2), c["sha256_map2"])
    img = texture(int(c["seed"]), iw, ih)
    assert np.array_equal(sha(img), c["sha256_image"]), "input generator drifted"
    assert np.array_equal(sha(oracle.remap_u8(img, m1, m2)), c["sha256_undistorted"])


def test_border_and_degenerate_maps_live(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (12, 16)).astype(np.uint8)
    m1 = np.zeros((8, 8, 2), np.int16)
    m1[..., 0] = rng.integers(-3, 20, (8, 8))
    m1[..., 1] = rng.integers(-3, 14, (8, 8))
    m2 = rng.integers(0, 1024, (8, 8)).astype(np.uint16)
    assert np.array_equal(cv2.remap(img, m1, m2, cv2.INTER_LINEAR), oracle.remap_u8(img, m1, m2))
    # identity map: exact copy; half-pixel map: rounding of (a + b + 1) >> 1 style averages
    ident = np.stack(np.meshgrid(np.arange(16), np.arange(12)), -1).astype(np.int16)
    assert np.array_equal(oracle.remap_u8(img, ident, np.zeros((12, 16), np.uint16)), img)
    half = np.full((12, 16), 16 * 32 + 16, np.uint16)
    assert np.array_equal(cv2.remap(img, ident, half, cv2.INTER_LINEAR), oracle.remap_u8(img, ident, half))
    sat = np.full((4, 4), 255, np.uint8)  # all-255 source never overflows the 15-bit weights
    idm = np.stack(np.meshgrid(np.arange(3), np.arange(3)), -1).astype(np.int16)
    assert (oracle.remap_u8(sat, idm, np.full((3, 3), 1023, np.uint16)) == 255).all()


def test_live_opencv_random_calibrations(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(4):
        iw, ih, ow, oh = 200, 150, 160, 128
        K = (150 + 40 * rng.random(), 150 + 40 * rng.random(), 99.5 + 6 * rng.standard_normal(), 74.5 + 6 * rng.standard_normal())
        dist = (-0.3 * rng.random(), 0.2 * rng.random(), 0.003 * rng.standard_normal(), 0.003 * rng.standard_normal(), 0.05 * rng.standard_normal())
        Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]])
        Kn, _ = cv2.getOptimalNewCameraMatrix(Km, np.array(dist), (iw, ih), float(rng.random()), (ow, oh))
        m1, m2 = cv2.initUndistortRectifyMap(Km, np.array(dist), None, Kn, (ow, oh), cv2.CV_16SC2)
        a1, a2 = oracle.init_undistort_rectify_map(K, dist, (Kn[0, 0], Kn[1, 1], Kn[0, 2], Kn[1, 2]), ow, oh)
        assert np.array_equal(a1, m1) and np.array_equal(a2, m2)
        img = texture(3, iw, ih)
        assert np.array_equal(cv2.remap(img, m1, m2, cv2.INTER_LINEAR), oracle.remap_u8(img, m1, m2))
