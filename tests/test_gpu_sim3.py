"""GPU parity for Sim3Tracker::trackFrameSim3 (B8-B11) against the oracle, through the C ABI.

Tolerances (north_star): per-iteration residual <= 1e-4 relative while the accept/reject sequences agree; final
Sim3 <= 1e-5 (translation in scene units, rotation in rad, scale) against the oracle's order-independent
(fp64-accumulator) mode, and no farther from the fp32 sequential mode than that mode is from the exact one.
"""
import numpy as np
import pytest

from common import make_sim3_pair, quat_angle

pytestmark = pytest.mark.gpu

RES_RTOL = 1e-4
POSE_TOL = 1e-5


def _gpu(lsd, d, w, h):
    ctx = lsd.Context(w, h, d["pr"]["K"])
    kf = ctx.create_frame(d["kf_img"], 0)
    fr = ctx.create_frame(d["fr_img"], 1)
    kf.set_idepth(d["idepth"], d["var"])
    fr.set_idepth(d["fr_idepth"], d["fr_var"])
    ref = ctx.create_refs([kf])[0]
    return ctx, kf, fr, ref


def _prefix(a, b):
    n = 0
    for x, y in zip(a, b):
        if (x[0], x[1]) != (y[0], y[1]):
            break
        n += 1
    return n


@pytest.mark.parametrize("seed,wh,c,levels", [(81, (640, 480), 1.0, (4, 1)), (82, (640, 480), 1.04, (4, 1)),
                                               (83, (320, 240), 0.95, (4, 3)), (84, (320, 240), 1.0, (2, 2)),
                                               (85, (320, 240), 1.02, (1, 1)),
                                               (86, (1280, 960), 1.03, (4, 1)), (87, (1280, 960), 0.97, (4, 3))])
def test_track_frame_sim3_matches_oracle(lsd, oracle, seed, wh, c, levels):
    w, h = wh
    from lsd_b200 import synth
    d = make_sim3_pair(oracle, seed, w, h, c=c, K=synth.d2_K() if wh == (1280, 960) else None)  # configs[4]: d2 intrinsics
    ctx, kf, fr, ref = _gpu(lsd, d, w, h)
    init = d["gt8"].copy()
    init[4:7] += [0.004, -0.003, 0.002]
    init[7] = 1.0
    g, gtrace = ctx.sim3_track(ref, fr, init, levels[0], levels[1], want_trace=True)
    e, etrace = oracle.sim3_track(d["oref"], d["ofr"], init, levels[0], levels[1], 2)
    gp, ep = np.array(g.frameToRef), np.array(e.frameToRef)
    agree = _prefix(gtrace, etrace)
    assert agree >= 2
    for k in range(agree):
        assert gtrace[k][4] == etrace[k][4], f"buf_warped_size at evaluation {k}"
        assert abs(gtrace[k][2] - etrace[k][2]) <= RES_RTOL * abs(etrace[k][2]), f"residual at evaluation {k}"
    assert g.diverged == e.diverged == 0
    assert np.linalg.norm(gp[4:7] - ep[4:7]) <= POSE_TOL, (gp, ep, agree, len(gtrace), len(etrace))
    assert quat_angle(gp[:4], ep[:4]) <= POSE_TOL
    assert abs(gp[7] - ep[7]) <= POSE_TOL
    if agree == len(etrace) == len(gtrace):
        assert np.isclose(g.lastResidual, e.lastResidual, rtol=RES_RTOL)
        assert np.isclose(g.lastDepthResidual, e.lastDepthResidual, rtol=RES_RTOL)
        assert np.isclose(g.lastPhotometricResidual, e.lastPhotometricResidual, rtol=RES_RTOL)
        assert np.isclose(g.pointUsage, e.pointUsage, rtol=1e-5)
        H, He = np.array(g.lastSim3Hessian), np.array(e.hessian)
        assert np.allclose(H, He, rtol=2e-4, atol=2e-5 * np.abs(He).max())
    o, _ = oracle.sim3_track(d["oref"], d["ofr"], init, levels[0], levels[1], 0)
    op = np.array(o.frameToRef)
    assert np.linalg.norm(gp[4:7] - op[4:7]) <= np.linalg.norm(op[4:7] - ep[4:7]) + POSE_TOL
    assert abs(gp[7] - op[7]) <= abs(op[7] - ep[7]) + POSE_TOL
    ctx.close()


def test_sim3_batch_is_deterministic_and_composition_independent(lsd, oracle):
    w, h = 320, 240
    ds = [make_sim3_pair(oracle, 90 + i, w, h, c=1.0 + 0.01 * i) for i in range(5)]
    ctx = lsd.Context(w, h, ds[0]["pr"]["K"])
    kfs = ctx.create_frames([d["kf_img"] for d in ds])
    frs = ctx.create_frames([d["fr_img"] for d in ds])
    for k, f, d in zip(kfs, frs, ds):
        k.set_idepth(d["idepth"], d["var"])
        f.set_idepth(d["fr_idepth"], d["fr_var"])
    refs = ctx.create_refs(kfs)
    inits = np.array([d["gt8"] for d in ds])
    inits[:, 7] = 1.0
    r1 = ctx.sim3_track_batch(refs, frs, inits)
    r2 = ctx.sim3_track_batch(refs, frs, inits)
    p1 = np.array([list(r.frameToRef) for r in r1])
    assert np.array_equal(p1, np.array([list(r.frameToRef) for r in r2]))
    for i in range(5):
        rs = ctx.sim3_track(refs[i], frs[i], inits[i])
        assert np.array_equal(np.array(rs.frameToRef), p1[i])
        assert abs(p1[i][7] - ds[i]["gt8"][7]) < 3e-3
    ctx.close()


def test_chained_stages_equal_separate_calls(lsd, oracle):
    """lsd_sim3_track_stages_batch ([4,3] -> [2] -> [1], SlamSystem::tryTrackSim3's schedule, inside one launch) returns for every
    stage exactly what the corresponding separate lsd_sim3_track_batch call returns when it is started from the previous call's
    result; a track that diverges stops and reports diverged for its remaining stages."""
    w, h = 640, 480
    ds = [make_sim3_pair(oracle, 120 + i, w, h, c=1.0 + 0.01 * i) for i in range(4)]
    ctx = lsd.Context(w, h, ds[0]["pr"]["K"])
    kfs = ctx.create_frames([d["kf_img"] for d in ds])
    frs = ctx.create_frames([d["fr_img"] for d in ds])
    for k, f, d in zip(kfs, frs, ds):
        k.set_idepth(d["idepth"], d["var"])
        f.set_idepth(d["fr_idepth"], d["fr_var"])
    refs = ctx.create_refs(kfs)
    inits = np.array([d["gt8"] for d in ds])
    inits[:, 4:7] += [0.004, -0.003, 0.002]
    inits[:, 7] = 1.0
    s_ = np.sin(np.pi / 4)
    inits[3] = [0, s_, 0, s_, 0, 0, 0, 1.0]  # looks 90 degrees away: diverges in the first stage
    stages = ((4, 3), (2, 2), (1, 1))
    chained = ctx.sim3_track_stages_batch(refs, frs, inits, stages)
    cur = inits.copy()
    for k, (ls, le) in enumerate(stages):
        sep = ctx.sim3_track_batch(refs[:3], frs[:3], cur[:3], ls, le)
        for i in range(3):
            a, b = chained[k][i], sep[i]
            assert list(a.frameToRef) == list(b.frameToRef), (k, i)
            assert list(a.lastSim3Hessian) == list(b.lastSim3Hessian)
            assert (a.lastResidual, a.pointUsage, a.affine_a, a.affine_b, a.diverged) == (b.lastResidual, b.pointUsage, b.affine_a, b.affine_b, b.diverged)
            assert list(a.numResidualCalls) == list(b.numResidualCalls) and list(a.numWarpUpdateCalls) == list(b.numWarpUpdateCalls)
            cur[i] = list(b.frameToRef)
        assert chained[k][3].diverged == 1 and list(chained[k][3].frameToRef) == [0, 0, 0, 1, 0, 0, 0, 1]
    # and the chain ends close to the truth (scale included)
    for i in range(3):
        assert abs(chained[2][i].frameToRef[7] - ds[i]["gt8"][7]) < 3e-3
    ctx.close()


def test_sim3_without_depth_residuals_takes_a_finite_step(lsd, oracle):
    """No warped point lands on a frame pixel with a depth hypothesis (numTermsD == 0): the scale row and column of the
    7x7 system are zero.  Upstream's Eigen LDLT applies the pseudo-inverse of D there (inc[6] = 0, finite step); an
    unguarded LDL^T would return NaN and the track would collapse to identity.  Device and oracle must both keep tracking."""
    w, h = 320, 240
    d = make_sim3_pair(oracle, 96, w, h)
    none_id = np.full((h, w), -1.0, np.float32)
    d["ofr"].set_idepth(none_id, none_id)
    ctx, kf, fr, ref = _gpu(lsd, d, w, h)
    fr.set_idepth(none_id, none_id)
    init = d["gt8"].copy()
    init[4:7] += [0.004, -0.003, 0.002]
    init[7] = 1.0
    g, gtrace = ctx.sim3_track(ref, fr, init, 4, 2, want_trace=True)
    e, etrace = oracle.sim3_track(d["oref"], d["ofr"], init, 4, 2, 2)
    gp, ep = np.array(g.frameToRef), np.array(e.frameToRef)
    assert np.isfinite(gp).all() and np.isfinite(ep).all()
    assert g.diverged == 0 and e.diverged == 0
    assert len(gtrace) > 6 and any(t[1] == 1 for t in gtrace), "steps must be taken and accepted"
    assert gp[7] == 1.0 and ep[7] == 1.0, "no depth residual: the scale must not move"
    assert np.linalg.norm(gp[4:7] - ep[4:7]) <= POSE_TOL and quat_angle(gp[:4], ep[:4]) <= POSE_TOL
    assert np.linalg.norm(gp[4:7] - init[4:7]) > 1e-4, "the pose must have moved"
    ctx.close()


def test_sim3_divergence(lsd, oracle):
    w, h = 320, 240
    d = make_sim3_pair(oracle, 95, w, h)
    ctx, kf, fr, ref = _gpu(lsd, d, w, h)
    s = np.sin(np.pi / 4)
    g = ctx.sim3_track(ref, fr, np.array([0, s, 0, s, 0, 0, 0, 1.0]))
    assert g.diverged == 1 and list(g.frameToRef) == [0, 0, 0, 1, 0, 0, 0, 1]
    ctx.close()
