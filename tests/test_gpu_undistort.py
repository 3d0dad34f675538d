"""GPU parity of the undistortion through the C ABI: maps and undistorted images bit-exact against
OpenCV's own output (tests/golden/undistort.npz), and the fused undistort -> Frame path equal to undistort-then-ingest."""
import numpy as np
import pytest

from test_oracle_undistort import FULL, G, SMALL, sha, texture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", SMALL)
def test_small_cases_bit_exact(lsd, name):
    c = G[name]
    h, w = c["map2"].shape
    iw, ih = (int(v) for v in c["in_wh"])
    ctx = lsd.Context(w, h, tuple(float(v) for v in c["Kout"]))
    und = lsd.Undistorter(ctx, iw, ih, K=c["K"], dist=c["dist"], K_out=c["Kout"])
    m1, m2 = und.maps()
    assert np.array_equal(m1, c["map1"]) and np.array_equal(m2, c["map2"]), "initUndistortRectifyMap"
    assert np.array_equal(und.undistort(c["image"]), c["undistorted"]), "remap"
    und2 = lsd.Undistorter(ctx, iw, ih, maps=(c["map1"], c["map2"]))  # maps handed over by an existing OpenCV undistorter
    assert np.array_equal(und2.undistort(c["image"]), c["undistorted"])
    und.close()
    und2.close()
    ctx.close()


@pytest.mark.parametrize("name", FULL)
def test_d2_camera_full_size_digests(lsd, name):
    """1920x1080 d2_camera.xml calibration -> 1280x960 / 640x480 (BASELINE configs[4] / configs[0] shapes)."""
    c = G[name]
    ow, oh = (int(v) for v in c["out_wh"])
    iw, ih = (int(v) for v in c["in_wh"])
    ctx = lsd.Context(ow, oh, tuple(float(v) for v in c["Kout"]))
    und = lsd.Undistorter(ctx, iw, ih, K=c["K"], dist=c["dist"], K_out=c["Kout"])
    m1, m2 = und.maps()
    assert np.array_equal(sha(m1), c["sha256_map1"]) and np.array_equal(sha(m2), c["sha256_map2"])
    img = texture(int(c["seed"]), iw, ih)
    assert np.array_equal(sha(und.undistort(img)), c["sha256_undistorted"])
    und.close()
    ctx.close()


def test_fused_frame_creation_equals_undistort_then_ingest(lsd, oracle):
    c = G["d2_small_crop"]
    h, w = c["map2"].shape
    iw, ih = (int(v) for v in c["in_wh"])
    ctx = lsd.Context(w, h, tuple(float(v) for v in c["Kout"]))
    und = lsd.Undistorter(ctx, iw, ih, maps=(c["map1"], c["map2"]))
    imgs = [texture(20 + i, iw, ih) for i in range(3)]
    frames, shown = und.create_frames(imgs, ids=[5, 6, 7], flags=lsd.BUILD_MAXGRAD0, want_undistorted=True)
    for im, f, s in zip(imgs, frames, shown):
        want = oracle.remap_u8(im, c["map1"], c["map2"])
        assert np.array_equal(s, want)
        g = ctx.create_frame(want, 99, flags=lsd.BUILD_MAXGRAD0)
        for l in range(5):
            assert np.array_equal(f.image(l), g.image(l))
        assert np.array_equal(f.maxGradients(0), g.maxGradients(0))
        g.release()
    with pytest.raises(AssertionError):
        und.undistort(np.zeros((ih + 1, iw), np.uint8))
    und.close()
    ctx.close()
