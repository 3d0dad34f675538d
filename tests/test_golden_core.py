"""Regression fixtures of the core restatement (tests/golden/core_regression.npz, scripts/make_golden_core.py).

NOT a parity pin -- the reference has no golden data for this path.  The CPU test keeps the oracle
from drifting between rounds; the GPU test checks the CUDA path against the committed numbers (in addition to the live
oracle comparisons of tests/test_gpu_*.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "scripts"))
G = dict(np.load(os.path.join(HERE, "golden", "core_regression.npz")))


def test_oracle_reproduces_committed_vectors(oracle):
    import make_golden_core as M
    now = M.build()
    assert sorted(now) == sorted(G)
    for k, v in G.items():
        if k.startswith(("pair/sha_", "depth/sha_")) or v.dtype.kind in "iu":
            assert np.array_equal(now[k], v), k
        else:
            assert np.allclose(now[k], v, rtol=1e-6, atol=1e-9), k


@pytest.mark.gpu
def test_device_matches_committed_vectors(lsd, oracle):
    import make_golden_core as M
    from common import hyp_from_idepth, make_oracle_depth_scene, make_oracle_pair
    w, h = 160, 112
    K = tuple(float(k) for k in G["pair/K"])
    d = make_oracle_pair(5, M.PAIR_W, M.PAIR_H)
    assert np.array_equal(M.sha(d["kf_img"]), G["pair/sha_kf_img"]) and np.array_equal(M.sha(d["idepth"]), G["pair/sha_idepth"])
    ctx = lsd.Context(M.PAIR_W, M.PAIR_H, K)
    kf = ctx.create_frame(d["kf_img"], 10, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    fr = ctx.create_frame(d["fr_img"], 11)
    for l in range(5):
        assert np.array_equal(M.sha(kf.image(l)), G[f"pair/sha_image_L{l}"]), f"image L{l}"
        assert np.array_equal(M.sha(kf.gradients(l)), G[f"pair/sha_gradients_L{l}"]), f"gradients L{l}"
    assert np.array_equal(M.sha(kf.maxGradients(0)), G["pair/sha_maxgrad_L0"])
    assert kf.num_mappable_pixels() == int(G["pair/num_mappable"])
    kf.set_idepth(d["idepth"], d["var"])
    ref = ctx.create_refs([kf])[0]
    assert [ref.num_data(l) for l in (1, 2, 3, 4)] == G["pair/numData"].tolist()
    init = np.array([0, 0, 0, 1, 0, 0, 0.0])
    res, trace = ctx.se3_track(ref, fr, init, want_trace=True)
    assert np.abs(np.array(res.frameToRef) - G["pair/se3_frameToRef"]).max() <= 1e-5
    gt = G["pair/se3_trace"]
    assert len(trace) == len(gt)
    for t, g in zip(trace, gt):
        assert (t[0], t[1], t[4]) == (int(g[0]), int(g[1]), int(g[3])) and abs(t[2] - g[2]) <= 1e-4 * abs(g[2]) + 1e-6
    pres = ctx.se3_track_permaref_batch([ref], [fr], [init])[0]
    assert np.abs(np.array(pres.frameToRef) - G["pair/permaref_refToFrame"]).max() <= 1e-5
    ov = ctx.check_permaref_overlap_batch([ref], [G["pair/permaref_refToFrame"]])[0]
    assert abs(ov - float(G["pair/permaref_overlap"])) <= 1e-5
    ctx.close()
    # DepthMap: canonical hypothesis maps bit for bit
    sc = make_oracle_depth_scene(4, w, h, n_refs=3)
    ctx = lsd.Context(w, h, sc["K"])
    dkf = ctx.create_frame(sc["kf_img"], 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    refs = []
    for i, r in enumerate(sc["refs"]):
        f = ctx.create_frame(r["img"], 1001 + i, flags=lsd.BUILD_MAXGRAD0)
        f.set_tracking_meta(1000, r["toParent"], 1.0)
        refs.append(f)
    dm = ctx.create_depthmap()
    dm.initializeFromMap(dkf, hyp_from_idepth(sc["idepth"], sc["var"]))
    dm.updateKeyframe(refs)
    m1 = dm.read()
    assert int(m1["isValid"].sum()) == int(G["depth/valid_after_update"])
    assert np.array_equal(M.sha(M.canonical_map(m1)), G["depth/sha_map_after_update"])
    dm.createKeyFrame(refs[-1])
    m2 = dm.read()
    assert int(m2["isValid"].sum()) == int(G["depth/valid_after_create"])
    assert np.array_equal(M.sha(M.canonical_map(m2)), G["depth/sha_map_after_create"])
    dm.destroy()
    ctx.close()
