"""CPU tests of the DepthMap oracle against analytic ground truth (SURVEY.md 8c sanity anchors (3)).

The reference holds no golden vector for this path (PARITY UNPINNED), so the restatement of
DepthMap::{observeDepth, doLineStereo, propagateDepth, regularizeDepthMap, regularizeDepthMapFillHoles,
createKeyFrame} is anchored on rendered scenes whose inverse depth is known exactly.
"""
import numpy as np

from common import hyp_from_idepth, make_oracle_depth_scene

W, H = 320, 240


def _rel_err(m, gt, sel):
    return np.abs(m["idepth"][sel] - gt[sel]) / gt[sel]


def test_line_stereo_creates_depth_close_to_ground_truth(oracle):
    """observeDepthCreate on an empty map: epipolar search over the full range [0, 1/MIN_DEPTH] finds GT idepth."""
    d = make_oracle_depth_scene(3, W, H, n_refs=6, step=0.02)
    dm = oracle.DepthMap(W, H, d["K"])
    dm.init_map(d["okf"], hyp_from_idepth(np.zeros((H, W), np.float32), -np.ones((H, W), np.float32)))
    dm.prepare([r["of"] for r in d["refs"][-1:]])   # one frame, 12 cm baseline
    dm.stage(oracle.STAGE_OBSERVE)
    m = dm.read()
    valid = m["isValid"] > 0
    assert valid.sum() > 0.05 * W * H
    err = _rel_err(m, d["gt_idepth"], valid)
    assert np.median(err) < 0.02
    assert np.mean(err < 0.1) > 0.9
    assert np.all(m["validity_counter"][valid] == 5) and np.all(m["idepth_smoothed"][valid] == -1)
    assert np.all(m["idepth_var"][valid] <= 0.25) and np.all(m["idepth_var"][valid] > 0)
    # nothing outside the [3, w-3) x [3, h-3) window or below the gradient threshold is ever created
    assert not valid[:3].any() and not valid[-3:].any() and not valid[:, :3].any() and not valid[:, -3:].any()
    assert not valid[d["maxgrad"] < 5].any()


def test_update_keyframe_reduces_error_and_variance(oracle):
    d = make_oracle_depth_scene(4, W, H, n_refs=10, noise=0.05, var=0.01)
    dm = oracle.DepthMap(W, H, d["K"])
    m0 = hyp_from_idepth(d["idepth"], d["var"])
    dm.init_map(d["okf"], m0)
    v0 = m0["isValid"] > 0
    e0 = np.median(_rel_err(m0, d["gt_idepth"], v0))
    # the live pipeline maps every tracked frame: one observation per pixel per call (each pixel uses ONE reference frame)
    dm.update_keyframe([d["refs"][0]["of"]])
    idp = d["okf"].get(oracle.IDEPTH, 0).copy()
    m1 = dm.read()
    for r in d["refs"][1:]:
        dm.update_keyframe([r["of"]])
    m2 = dm.read()
    v1 = m2["isValid"] > 0
    both = v0 & v1
    assert both.sum() > 0.8 * v0.sum()
    e1 = np.median(_rel_err(m2, d["gt_idepth"], both))
    assert e1 < 0.7 * e0, (e0, e1)
    upd = both & (m2["idepth_var"] < np.float32(0.01))  # pixels whose gradient runs along the epipolar lines got fused
    assert upd.sum() > 0.25 * both.sum()
    assert np.median(_rel_err(m2, d["gt_idepth"], upd)) < 0.5 * e0
    v1 = m1["isValid"] > 0
    # Frame::setDepth ran (depthHasBeenUpdatedFlag was false): level-0 idepth == smoothed idepth of valid pixels
    ok = v1 & (m1["idepth_smoothed"] >= -0.05)
    assert np.array_equal(idp[ok], m1["idepth_smoothed"][ok]) and np.all(idp[~ok] == -1)
    assert oracle.frame_counters(d["okf"])[1] == 10  # numMappedOnThis


def test_regularize_and_fill_holes_invariants(oracle):
    d = make_oracle_depth_scene(5, W, H, n_refs=2)
    dm = oracle.DepthMap(W, H, d["K"])
    m0 = hyp_from_idepth(d["idepth"], d["var"], validity=20)
    # punch holes into textured regions
    rng = np.random.default_rng(0)
    holes = (rng.random((H, W)) < 0.2) & (m0["isValid"] > 0)
    m0["isValid"][holes] = 0
    dm.init_map(d["okf"], m0)
    dm.stage(oracle.STAGE_FILL_HOLES)
    m1 = dm.read()
    was = m0["isValid"] > 0
    assert np.array_equal(m1[was], m0[was]), "fillHoles must not touch valid pixels"
    filled = (m1["isValid"] > 0) & ~was
    assert filled.sum() > 0.5 * holes[3:-3, 3:-3].sum()
    assert np.all(m1["validity_counter"][filled] == 0) and np.all(m1["idepth_var"][filled] == np.float32(0.125))
    assert np.median(np.abs(m1["idepth"][filled] - d["gt_idepth"][filled]) / d["gt_idepth"][filled]) < 0.1
    dm.stage(oracle.STAGE_REGULARIZE, 0, 24)
    m2 = dm.read()
    v2 = m2["isValid"] > 0
    inner = np.zeros_like(v2)
    inner[2:-2, 2:-2] = True
    assert np.all(m2["idepth_var_smoothed"][v2 & inner] > 0)
    # smoothing is a convex combination of neighbours within one sigma: stays within the neighbourhood range
    assert np.median(np.abs(m2["idepth_smoothed"][v2 & inner] - d["gt_idepth"][v2 & inner]) / d["gt_idepth"][v2 & inner]) < 0.05


def test_create_keyframe_propagates_depth_to_new_view(oracle):
    d = make_oracle_depth_scene(6, W, H, n_refs=10, noise=0.0, var=0.001, with_mask=True)
    dm = oracle.DepthMap(W, H, d["K"])
    dm.init_map(d["okf"], hyp_from_idepth(d["idepth"], d["var"], validity=30))
    new = d["refs"][-1]
    dm.create_keyframe(new["of"])
    m = dm.read()
    v = m["isValid"] > 0
    assert v.sum() > 0.5 * (d["var"] > 0).sum()
    s = dm.last_rescale()
    gt_new = 1.0 / new["depth"]
    err = np.abs(m["idepth"][v] / s - gt_new[v]) / gt_new[v]
    assert np.median(err) < 0.01, np.median(err)
    assert abs(np.mean(m["idepth_smoothed"][v]) - 1.0) < 1e-3  # mean inverse depth normalised to one
    pose = oracle.frame_pose(new["of"])
    assert abs(pose[7] - s) < 1e-6 and np.allclose(pose[:7], new["toParent"][:7], atol=1e-9)
    idp = new["of"].get(oracle.IDEPTH, 0)
    assert np.array_equal(idp[v], m["idepth_smoothed"][v])


def test_threading_does_not_change_results(oracle):
    d = make_oracle_depth_scene(7, W, H, n_refs=4)
    out = []
    for threads in (1, 4):
        dd = make_oracle_depth_scene(7, W, H, n_refs=4)
        dm = oracle.DepthMap(W, H, dd["K"], threads=threads)
        dm.init_map(dd["okf"], hyp_from_idepth(dd["idepth"], dd["var"]))
        dm.update_keyframe([r["of"] for r in dd["refs"]])
        out.append(dm.read())
    assert out[0].tobytes() == out[1].tobytes()


def test_debug_plot_is_rgb_rainbow(oracle):
    d = make_oracle_depth_scene(8, 64, 48, n_refs=1)
    dm = oracle.DepthMap(64, 48, d["K"])
    dm.init_map(d["okf"], hyp_from_idepth(d["idepth"], d["var"]))
    rgb = dm.debug_rgb()
    v = d["var"] > 0
    assert rgb.shape == (48, 64, 3)
    assert np.all(rgb[~v][:, 0] == rgb[~v][:, 1])  # grey where there is no hypothesis
    idv = d["idepth"][v]
    exp_r = 255 - np.minimum(255, np.abs(0 - idv) * 255).astype(np.uint8)
    assert np.array_equal(rgb[v][:, 0], exp_r)
