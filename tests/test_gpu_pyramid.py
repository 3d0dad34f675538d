"""GPU parity: Frame pyramids / masks / counts must be BIT-EXACT against the oracle (north_star)."""
import numpy as np
import pytest

from common import make_oracle_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wh", [(640, 480), (320, 240), (64, 48), (1280, 960)])
def test_image_gradient_maxgrad_bit_exact(lsd, oracle, synth, wh):
    w, h = wh
    pr = synth.make_pair(21, w, h)
    img = pr["kf_img"].numpy()
    of = oracle.Frame(0, img, pr["K"])
    of.build_pyramids()
    ctx = lsd.Context(w, h, pr["K"])
    gf = ctx.create_frame(img, 0, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    for l in range(5):
        assert np.array_equal(gf.image(l), of.get(oracle.IMAGE, l)), f"image L{l}"
        assert np.array_equal(gf.gradients(l), of.get(oracle.GRADIENTS, l)), f"gradients L{l}"
    assert np.array_equal(gf.maxGradients(0), of.get(oracle.MAXGRAD, 0))
    assert gf.num_mappable_pixels() == of.num_mappable()
    gf.release()
    ctx.close()


def test_lazy_planes_match_eager(lsd, synth):
    pr = synth.make_pair(22, 320, 240)
    img = pr["kf_img"].numpy()
    ctx = lsd.Context(320, 240, pr["K"])
    a = ctx.create_frame(img, 0, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    b = ctx.create_frame(img, 1, flags=lsd.BUILD_TRACKING)
    assert np.array_equal(a.maxGradients(0), b.maxGradients(0))
    assert np.array_equal(a.gradients(0), b.gradients(0))
    assert a.num_mappable_pixels() == b.num_mappable_pixels()
    ctx.close()


def test_batch_ingest_equals_single(lsd, synth):
    w, h = 320, 240
    imgs = [synth.make_pair(30 + s, w, h)["fr_img"].numpy() for s in range(5)]
    K = synth.default_K(w, h)
    ctx = lsd.Context(w, h, K)
    batch = ctx.create_frames(imgs, list(range(5)))
    for i, im in enumerate(imgs):
        single = ctx.create_frame(im, 100 + i)
        for l in range(5):
            assert np.array_equal(batch[i].image(l), single.image(l))
            if l >= 1:
                assert np.array_equal(batch[i].gradients(l), single.gradients(l))
        single.release()
    ctx.close()


def test_extreme_images(lsd, oracle):
    """all-0, all-255 and a checkerboard (max gradients) -- edge cases of the exact-sum argument."""
    w, h = 64, 48
    K = (52.5, 52.5, 31.5, 23.5)
    ctx = lsd.Context(w, h, K)
    yy, xx = np.mgrid[0:h, 0:w]
    for img in [np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8), (((xx + yy) & 1) * 255).astype(np.uint8)]:
        of = oracle.Frame(0, img, K)
        of.build_pyramids()
        gf = ctx.create_frame(img, 0, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
        for l in range(5):
            assert np.array_equal(gf.image(l), of.get(oracle.IMAGE, l))
            assert np.array_equal(gf.gradients(l), of.get(oracle.GRADIENTS, l))
        assert np.array_equal(gf.maxGradients(0), of.get(oracle.MAXGRAD, 0))
        assert gf.num_mappable_pixels() == of.num_mappable()
        gf.release()
    ctx.close()


@pytest.mark.parametrize("wh", [(640, 480), (64, 48)])
def test_idepth_pyramid_and_pointcloud(lsd, oracle, wh):
    w, h = wh
    d = make_oracle_pair(23, w, h)
    ctx = lsd.Context(w, h, d["pr"]["K"])
    kf = ctx.create_frame(d["kf_img"], 0)
    kf.set_idepth(d["idepth"], d["var"])
    ref = ctx.create_refs([kf])[0]
    for l in range(1, 5):
        assert np.array_equal(kf.idepth(l), d["okf"].get(oracle.IDEPTH, l)), f"idepth L{l}"
        assert np.array_equal(kf.idepthVar(l), d["okf"].get(oracle.IDEPTHVAR, l)), f"idepthVar L{l}"
        assert ref.num_data(l) == d["oref"].num(l), f"numData L{l}"
        gpos, ggrad, gcv, gidx = ref.read(l)
        opos, ograd, ocv, oidx = d["oref"].get(l)
        # same set of points, different emission order (row-major vs column-major): sort both by idx
        go, oo = np.argsort(gidx), np.argsort(oidx)
        assert np.array_equal(gidx[go], oidx[oo])
        assert np.array_equal(gpos[go], opos[oo])
        assert np.array_equal(ggrad[go], ograd[oo])
        assert np.array_equal(gcv[go], ocv[oo])
        assert np.all(np.diff(gidx) > 0)  # row-major, strictly increasing
    ctx.close()


def test_depth_from_gt(lsd, oracle, synth):
    w, h = 320, 240
    pr = synth.make_pair(24, w, h)
    depth = pr["kf_depth"].numpy().copy()
    depth[10:20, 30:50] = 0  # invalid GT
    depth[100, 100] = -1
    of = oracle.Frame(0, pr["kf_img"].numpy(), pr["K"])
    of.set_depth_gt(depth, 1.0)
    ctx = lsd.Context(w, h, pr["K"])
    gf = ctx.create_frame(pr["kf_img"].numpy(), 0)
    gf.set_depth_from_gt(depth, 1.0)
    for l in range(5):
        assert np.array_equal(gf.idepth(l), of.get(oracle.IDEPTH, l))
        assert np.array_equal(gf.idepthVar(l), of.get(oracle.IDEPTHVAR, l))
    ctx.close()
