"""GPU parity of the keyframe publish / VBO extraction through the C ABI: bit-exact against the vectors
recorded from the reference's own Keyframe::computeVbo (tests/golden/keyframe_vbo.npz) and against the pinned restatement
(oracle/keyframe.cpp) at BASELINE sizes."""
import numpy as np
import pytest

from test_oracle_keyframe import CASES, _random_case

pytestmark = pytest.mark.gpu


def _device_frame(ctx, lsd, idepth, var, image_u8, fid=0):
    f = ctx.create_frame(np.ascontiguousarray(image_u8, np.uint8), fid)
    f.set_idepth(idepth, var)
    return f


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_vectors_bit_exact(lsd, name):
    c = CASES[name]
    h, w = c["idepth"].shape
    ctx = lsd.Context(w, h, tuple(float(k) for k in c["K"]))
    f = _device_frame(ctx, lsd, c["idepth"], c["var"], c["image"])
    pts = f.publish_keyframe(0)
    assert pts.tobytes() == c["points"].tobytes(), "publishKeyframe pack"
    vtx = f.compute_vbo(float(c["scale"]))
    assert len(vtx) == c["vertices"].size // 16, "points"
    assert vtx.tobytes() == c["vertices"].tobytes(), "vertex buffer (order, positions, colours)"
    ctx.close()


@pytest.mark.parametrize("wh,valid,scale", [((640, 480), 0.97, 1.0), ((640, 480), 0.5, 12.0), ((1280, 960), 0.9, 0.7), ((64, 48), 0.0, 1.0),
                                            ((64, 48), 1.0, 1.0)])
def test_baseline_sizes_equal_restatement(lsd, oracle, wh, valid, scale):
    w, h = wh
    idp, var, img = _random_case(w + h, w, h, valid)
    K = (0.82 * w, 0.83 * w, w / 2 - 0.5, h / 2 - 0.5)
    ctx = lsd.Context(w, h, K)
    f = _device_frame(ctx, lsd, idp, var, img.astype(np.uint8))
    want = oracle.compute_vbo(oracle.publish_keyframe_pack(idp, var, img), np.array(K, np.float32), scale)
    got = f.compute_vbo(scale)
    assert len(got) == len(want)
    assert got.tobytes() == want.tobytes()
    if valid == 0.0:
        assert len(got) == 0
    # FMA-contracted variant (what a -march=native build of the reference computes)
    prm = ctx.default_vbo_params()
    prm.contractFma = 1
    want_fma = oracle.compute_vbo(oracle.publish_keyframe_pack(idp, var, img), np.array(K, np.float32), scale, contract_fma=True)
    assert f.compute_vbo(scale, params=prm).tobytes() == want_fma.tobytes()
    ctx.close()


def test_pyramid_level_and_batch(lsd, oracle):
    """publishLvl > 0 reads the idepth pyramid; a batch launch (blockIdx.y = keyframe) equals the single calls."""
    w, h = 320, 240
    K = (262.5, 262.5, 159.5, 119.5)
    ctx = lsd.Context(w, h, K)
    frames, singles = [], []
    for s in range(5):
        idp, var, img = _random_case(100 + s, w, h, 0.6 + 0.08 * s)
        frames.append(_device_frame(ctx, lsd, idp, var, img.astype(np.uint8), s))
    scales = [1.0, 0.5, 2.0, 9.0, 1.0]
    for lvl in (0, 1):
        singles = [f.compute_vbo(sc, level=lvl) for f, sc in zip(frames, scales)]
        pts, outs = ctx.compute_vbo_batch(frames, scales, level=lvl)
        for i in range(5):
            assert pts[i] == len(singles[i])
            assert outs[i].tobytes() == singles[i].tobytes()
        # level-l restatement from the device's own planes (pyramids are bit-exact, tests/test_gpu_pyramid.py)
        f = frames[2]
        Kl = np.array([K[0] / 2 ** lvl, K[1] / 2 ** lvl, (K[2] + 0.5) / 2 ** lvl - 0.5, (K[3] + 0.5) / 2 ** lvl - 0.5], np.float32)
        want = oracle.compute_vbo(oracle.publish_keyframe_pack(f.idepth(lvl), f.idepthVar(lvl), f.image(lvl)), Kl, scales[2])
        assert singles[2].tobytes() == want.tobytes()
        assert f.publish_keyframe(lvl).tobytes() == oracle.publish_keyframe_pack(f.idepth(lvl), f.idepthVar(lvl), f.image(lvl)).tobytes()
    ctx.close()


def test_errors_and_no_depth(lsd):
    w, h = 64, 48
    ctx = lsd.Context(w, h, (52.5, 52.5, 31.5, 23.5))
    f = ctx.create_frame(np.zeros((h, w), np.uint8), 0)
    assert not f.publish_keyframe(0).view(np.uint8).any()  # reference: warning + unfilled buffer
    with pytest.raises(lsd.LsdError):
        f.compute_vbo(1.0)
    f.set_idepth(np.ones((h, w), np.float32), np.full((h, w), 1e-4, np.float32))
    prm = ctx.default_vbo_params()
    prm.sparsifyFactor = 2
    with pytest.raises(lsd.LsdError):
        f.compute_vbo(1.0, params=prm)
    assert len(f.compute_vbo(1.0)) == (w - 2) * (h - 2)
    ctx.close()


def test_pipeline_keyframe_publishes(lsd, oracle, synth):
    """End of the hot path: a keyframe whose depth came from DepthMap (setDepth) publishes the same cloud as the oracle's."""
    from common import make_oracle_depth_scene, hyp_from_idepth
    w, h = 320, 240
    d = make_oracle_depth_scene(3, w, h, n_refs=3, var=1e-4, noise=0.002)  # converged map: passes computeVbo's variance gates
    ctx = lsd.Context(w, h, d["K"])
    kf = ctx.create_frame(d["kf_img"], 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    dm = ctx.create_depthmap()
    m0 = hyp_from_idepth(d["idepth"], d["var"])
    dm.initializeFromMap(kf, m0)
    dm.finalizeKeyFrame()
    odm = oracle.DepthMap(w, h, d["K"])
    odm.init_map(d["okf"], m0)
    odm.finalize()
    want = oracle.compute_vbo(oracle.publish_keyframe_pack(d["okf"].get(oracle.IDEPTH, 0), d["okf"].get(oracle.IDEPTHVAR, 0),
                                                           d["okf"].get(oracle.IMAGE, 0)), np.array(d["K"], np.float32), 1.0)
    got = kf.compute_vbo(1.0)
    assert len(want) > 100
    assert got.tobytes() == want.tobytes()
    dm.destroy()
    ctx.close()
