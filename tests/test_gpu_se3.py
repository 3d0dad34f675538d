"""GPU parity for the SE3 tracker (B3-B6) against the oracle, through the C ABI.

Tolerances (north_star): counts / masks bit-exact given identical pose inputs; per-iteration
residual <= 1e-4 relative while the accept/reject sequences agree; final SE3 <= 1e-5
(translation in scene units, rotation in rad) against BOTH oracle reduction orders.
"""
import numpy as np
import pytest

from common import make_oracle_pair, quat_angle

pytestmark = pytest.mark.gpu

RES_RTOL = 1e-4
POSE_TOL = 1e-5


def _gpu_pair(lsd, d, w, h):
    ctx = lsd.Context(w, h, d["pr"]["K"])
    kf = ctx.create_frame(d["kf_img"], 0)
    fr = ctx.create_frame(d["fr_img"], 1)
    kf.set_idepth(d["idepth"], d["var"])
    ref = ctx.create_refs([kf])[0]
    return ctx, kf, fr, ref


@pytest.mark.parametrize("level", [1, 2, 3, 4])
def test_fused_evaluation_matches_three_reference_passes(lsd, oracle, level):
    """One fused GPU evaluation == calcResidualAndBuffers + calcWeightsAndResidual + calculateWarpUpdate.

    EXACT (mode 2) is the oracle with fp64 accumulators: the order-independent value of every sum.
    The GPU (fp32 tree sums) must be within 1e-4 of it; the oracle's own fp32 modes (0 scalar,
    1 sse4-order) carry sequential-summation noise, so against them the bound is the triangle
    inequality |gpu - mode| <= |mode - exact| + tol.
    """
    w, h = 640, 480
    d = make_oracle_pair(41, w, h)
    ctx, kf, fr, ref = _gpu_pair(lsd, d, w, h)
    from test_oracle_tracking import _inv_pose7
    pose = _inv_pose7(d["pr"]["frameToRef"])  # refToFrame near the optimum
    for (a, b) in [(1.0, 0.0), (1.02, -1.5)]:
        gA, gb, gs = ctx.se3_eval(ref, fr, pose, level, a, b)
        eA, eb, es = oracle.se3_eval(d["oref"], d["ofr"], pose, level, a, b, 2)
        assert gs[2] == es[2] and gs[3] == es[3] and gs[4] == es[4], "bufSize / good / bad (bit-exact)"
        assert np.allclose(gs[[0, 1, 5, 6]], es[[0, 1, 5, 6]], rtol=RES_RTOL, atol=1e-6)
        # affine-lighting estimate sqrt((syy - sy^2/sw)/(sxx - sx^2/sw)) amplifies fp32 rounding of the sums
        # ~20x (a) / ~100x (b): compare a at 2e-4 and the fitted line a*mean+b (what enters residuals) tightly
        mean_c = float(d["okf"].get(oracle.IMAGE, level).mean())
        assert abs(gs[7] - es[7]) <= 1e-5 * abs(es[7])
        assert abs((gs[7] - es[7]) * mean_c + (gs[8] - es[8])) <= 1e-4
        assert np.allclose(gA, eA, rtol=1e-4, atol=1e-5 * np.abs(eA).max())
        assert np.allclose(gb, eb, rtol=1e-4, atol=1e-5 * np.abs(eb).max())
        for mode in (0, 1):
            oA, ob, os_ = oracle.se3_eval(d["oref"], d["ofr"], pose, level, a, b, mode)
            assert gs[2] == os_[2] and gs[3] == os_[3] and gs[4] == os_[4]
            for k in (0, 1, 5, 6, 7):
                assert abs(gs[k] - os_[k]) <= abs(os_[k] - es[k]) + RES_RTOL * abs(es[k]) + 1e-6, (k, gs[k], os_[k], es[k])
            assert np.all(np.abs(gA - oA) <= np.abs(oA - eA) + 1e-4 * np.abs(eA) + 1e-5 * np.abs(eA).max())
        if level == 1:
            assert np.array_equal(fr.refPixelWasGood(), d["ofr"].get(oracle.MASK, 1)), "refPixelWasGood mask"
    ctx.close()


def _agreeing_prefix(*traces):
    n = 0
    for rows in zip(*traces):
        if any((r[0], r[1]) != (rows[0][0], rows[0][1]) for r in rows):
            break
        n += 1
    return n


@pytest.mark.parametrize("seed,wh", [(41, (640, 480)), (42, (640, 480)), (43, (320, 240)), (44, (1280, 960))])
def test_track_frame_matches_oracle(lsd, oracle, seed, wh):
    """SE3Tracker::trackFrame through the C ABI vs the oracle.

    vs EXACT (fp64 accumulators): per-iteration residual <= 1e-4 relative while the accept/reject
    sequences agree, final SE3 <= 1e-5.  vs the fp32 SCALAR / SSE4-order modes: no farther than those
    modes are from EXACT, plus the same tolerance (their sequential fp32 sums are the noisier side).
    """
    w, h = wh
    d = make_oracle_pair(seed, w, h)
    ctx, kf, fr, ref = _gpu_pair(lsd, d, w, h)
    init = np.array([0, 0, 0, 1, 0, 0, 0.0])
    gres, gtrace = ctx.se3_track(ref, fr, init, want_trace=True)
    gpose = np.array(gres.frameToRef)
    gmask = fr.refPixelWasGood()
    eres, etrace = oracle.se3_track(d["oref"], d["ofr"], init, 2)
    emask = d["ofr"].get(oracle.MASK, 1).copy()
    epose = np.array(eres.frameToRef)

    agree = _agreeing_prefix(gtrace, etrace)
    assert agree >= 4
    for k in range(agree):
        g, e = gtrace[k], etrace[k]
        assert g[4] == e[4], f"buf_warped_size differs at evaluation {k}"
        assert abs(g[2] - e[2]) <= RES_RTOL * abs(e[2]), f"residual at evaluation {k}: {g[2]} vs {e[2]}"
    assert np.linalg.norm(gpose[4:] - epose[4:]) <= POSE_TOL, (gpose, epose, agree, len(gtrace), len(etrace))
    assert quat_angle(gpose[:4], epose[:4]) <= POSE_TOL
    assert gres.diverged == eres.diverged and gres.trackingWasGood == eres.trackingWasGood
    if agree == len(etrace) == len(gtrace):
        assert gres.lastGoodCount == eres.lastGoodCount and gres.lastBadCount == eres.lastBadCount
        assert list(gres.numResidualCalls) == list(eres.numResidualCalls)
        assert list(gres.numWarpUpdateCalls) == list(eres.numWarpUpdateCalls)
        assert np.isclose(gres.pointUsage, eres.pointUsage, rtol=1e-5)
        assert np.isclose(gres.lastResidual, eres.lastResidual, rtol=RES_RTOL)
        assert np.isclose(gres.initialTrackedResidual, eres.initialTrackedResidual, rtol=RES_RTOL)
        assert np.mean(gmask != emask) <= 1e-4  # identical up to isGood threshold ties

    for mode in (0, 1):
        ores, otrace = oracle.se3_track(d["oref"], d["ofr"], init, mode)
        opose = np.array(ores.frameToRef)
        n3 = _agreeing_prefix(gtrace, etrace, otrace)
        for k in range(n3):
            g, e, o = gtrace[k], etrace[k], otrace[k]
            assert abs(g[2] - o[2]) <= abs(o[2] - e[2]) + RES_RTOL * abs(e[2]), f"mode {mode} evaluation {k}"
        dt_oe = np.linalg.norm(opose[4:] - epose[4:])
        dr_oe = quat_angle(opose[:4], epose[:4])
        assert np.linalg.norm(gpose[4:] - opose[4:]) <= dt_oe + POSE_TOL
        assert quat_angle(gpose[:4], opose[:4]) <= dr_oe + POSE_TOL
    ctx.close()


def test_batch_equals_single_and_is_deterministic(lsd, oracle):
    w, h = 320, 240
    ds = [make_oracle_pair(50 + s, w, h) for s in range(6)]
    ctx = lsd.Context(w, h, ds[0]["pr"]["K"])
    kfs = ctx.create_frames([d["kf_img"] for d in ds])
    frs = ctx.create_frames([d["fr_img"] for d in ds])
    for k, d in zip(kfs, ds):
        k.set_idepth(d["idepth"], d["var"])
    refs = ctx.create_refs(kfs)
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (6, 1))
    r1 = ctx.se3_track_batch(refs, frs, inits)
    p1 = np.array([list(r.frameToRef) for r in r1])
    r2 = ctx.se3_track_batch(refs, frs, inits)
    p2 = np.array([list(r.frameToRef) for r in r2])
    assert np.array_equal(p1, p2), "batched tracking must be run-to-run deterministic"
    for recs in (1, 3, 8):  # work-item size is a scheduling knob only: results must not move by one bit
        ctx.set_se3_work_item_records(recs)
        r3 = ctx.se3_track_batch(refs, frs, inits)
        assert np.array_equal(np.array([list(r.frameToRef) for r in r3]), p1), f"result depends on work-item size {recs}"
    ctx.set_se3_work_item_records(0)
    r18 = ctx.se3_track_batch(refs * 3, frs * 3, np.tile(inits, (3, 1)))  # the same pairs inside a larger batch
    assert np.array_equal(np.array([list(r.frameToRef) for r in r18]), np.tile(p1, (3, 1)))
    for i in range(6):
        rs = ctx.se3_track(refs[i], frs[i], inits[i])
        assert np.array_equal(np.array(rs.frameToRef), p1[i]), "batch result must not depend on batch composition"
        ores, _ = oracle.se3_track(ds[i]["oref"], ds[i]["ofr"], inits[i], 2)
        assert np.linalg.norm(p1[i][4:] - np.array(ores.frameToRef)[4:]) <= POSE_TOL
    # Record size is the one knob that DEFINES the summation order (lsd_ctx_set_se3_record_points): for each value the
    # result is again independent of batch composition / scheduling, and it stays within the pose tolerance of the oracle.
    for pts in (1024, 256):
        ctx.set_se3_record_points(pts)
        rb = ctx.se3_track_batch(refs, frs, inits)
        pb = np.array([list(r.frameToRef) for r in rb])
        ctx.set_se3_work_item_records(4)
        assert np.array_equal(np.array([list(r.frameToRef) for r in ctx.se3_track_batch(refs, frs, inits)]), pb)
        ctx.set_se3_work_item_records(0)
        for i in range(6):
            assert np.array_equal(np.array(ctx.se3_track(refs[i], frs[i], inits[i]).frameToRef), pb[i])
            ores, _ = oracle.se3_track(ds[i]["oref"], ds[i]["ofr"], inits[i], 2)
            assert np.linalg.norm(pb[i][4:] - np.array(ores.frameToRef)[4:]) <= POSE_TOL
    ctx.set_se3_record_points(0)
    assert np.array_equal(np.array([list(r.frameToRef) for r in ctx.se3_track_batch(refs, frs, inits)]), p1)
    with pytest.raises(lsd.LsdError):
        ctx.set_se3_record_points(100)
    ctx.close()


def test_bench_batch_parity_64_pairs(lsd, oracle):
    """The first 64 pairs of bench.py's config-2 batch (same generator, same seeds) tracked in ONE batched launch vs the
    parity-build oracle in EXACT mode (fp64 accumulators), pair by pair:
      * the first LM evaluation (identical pose on both sides) has a bit-exact buf_warped_size and a residual <= 1e-4 rel;
      * pairs whose accept / reject sequences agree end within 1e-5 of the oracle (scene units, rad);
      * the accept / reject FLIP RATE (an LM step accepted on one side and rejected on the other, error ~ lastErr) is no
        higher than the oracle's own fp32 modes show against EXACT, and flipped pairs still agree to 1e-2 / within the
        triangle bound of the fp32 modes."""
    import torch

    import bench  # repo root is on sys.path (tests/conftest.py)
    n, w, h = 64, bench.W, bench.H
    K, kf, fr, idp, var, gt = bench.make_inputs(n, 0, torch.device("cuda", 0))
    kf_np, fr_np, id_np, var_np = kf.cpu().numpy(), fr.cpu().numpy(), idp.cpu().numpy(), var.cpu().numpy()
    ctx = lsd.Context(w, h, K)
    kfs = ctx.create_frames(list(kf_np))
    for k, a, b in zip(kfs, id_np, var_np):
        k.set_idepth(a, b)
    refs = ctx.create_refs(kfs)
    frs = ctx.create_frames(list(fr_np))
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (n, 1))
    gres, gtraces = ctx.se3_track_batch(refs, frs, inits, want_trace=True)
    flips = {0: 0, 1: 0, "gpu": 0}
    worst_same, worst_flip = 0.0, 0.0
    for i in range(n):
        okf, ofr = oracle.Frame(2 * i, kf_np[i], K), oracle.Frame(2 * i + 1, fr_np[i], K)
        okf.build_pyramids()
        ofr.build_pyramids()
        okf.set_idepth(id_np[i], var_np[i])
        oref = oracle.Ref(okf)
        eres, etrace = oracle.se3_track(oref, ofr, inits[i], 2)
        ep, gp = np.array(eres.frameToRef), np.array(gres[i].frameToRef)
        gtr = gtraces[i]
        assert gtr[0][4] == etrace[0][4], f"pair {i}: buf_warped_size of the first evaluation"
        assert abs(gtr[0][2] - etrace[0][2]) <= RES_RTOL * abs(etrace[0][2]), f"pair {i}: first residual"
        same = len(gtr) == len(etrace) and _agreeing_prefix(gtr, etrace) == len(etrace)
        dg = max(np.linalg.norm(gp[4:] - ep[4:]), quat_angle(gp[:4], ep[:4]))
        env = 0.0
        for mode in (0, 1):
            ores, otrace = oracle.se3_track(oref, ofr, inits[i], mode)
            op = np.array(ores.frameToRef)
            osame = len(otrace) == len(etrace) and _agreeing_prefix(otrace, etrace) == len(etrace)
            flips[mode] += 0 if osame else 1
            env = max(env, np.linalg.norm(op[4:] - ep[4:]), quat_angle(op[:4], ep[:4]))
        if same:
            worst_same = max(worst_same, dg)
            assert dg <= POSE_TOL, f"pair {i}: same accept/reject sequence but pose differs by {dg}"
            assert gres[i].lastGoodCount == eres.lastGoodCount and gres[i].lastBadCount == eres.lastBadCount
        else:
            flips["gpu"] += 1
            worst_flip = max(worst_flip, dg)
            assert dg <= max(1e-2, 3 * env), f"pair {i}: flipped LM sequence, pose differs by {dg} (fp32-mode envelope {env})"
        assert gres[i].diverged == eres.diverged and gres[i].trackingWasGood == eres.trackingWasGood
    print(f"bench-batch parity: {n} pairs, GPU-vs-EXACT flips {flips['gpu']}, SCALAR-vs-EXACT {flips[0]}, SSE4-vs-EXACT {flips[1]}, "
          f"worst pose diff (same sequence) {worst_same:.2e}, (flipped) {worst_flip:.2e}")
    assert flips["gpu"] <= max(flips[0], flips[1]) + max(3, n // 10), flips
    ctx.close()


def test_divergence_is_reported(lsd, oracle, synth):
    """A frame that does not overlap the keyframe at all: diverged, identity returned (upstream returns SE3())."""
    w, h = 320, 240
    d = make_oracle_pair(60, w, h)
    ctx, kf, fr, ref = _gpu_pair(lsd, d, w, h)
    # initial estimate looking 90 degrees away: every point projects outside
    s = np.sin(np.pi / 4)
    init = np.array([0, s, 0, s, 0, 0, 0.0])
    g = ctx.se3_track(ref, fr, init)
    o, _ = oracle.se3_track(d["oref"], d["ofr"], init, 0)
    assert g.diverged == 1 and o.diverged == 1
    assert g.trackingWasGood == 0
    assert list(g.frameToRef) == [0, 0, 0, 1, 0, 0, 0]
    ctx.close()


def test_host_image_entry_point(lsd, oracle):
    """lsd_se3_track_images_batch (H2D + pyramids + track) == create_frames + se3_track_batch."""
    w, h = 320, 240
    ds = [make_oracle_pair(70 + s, w, h) for s in range(4)]
    ctx = lsd.Context(w, h, ds[0]["pr"]["K"])
    kfs = ctx.create_frames([d["kf_img"] for d in ds])
    for k, d in zip(kfs, ds):
        k.set_idepth(d["idepth"], d["var"])
    refs = ctx.create_refs(kfs)
    frs = ctx.create_frames([d["fr_img"] for d in ds])
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (4, 1))
    a = ctx.se3_track_batch(refs, frs, inits)
    imgs = [np.ascontiguousarray(d["fr_img"]) for d in ds]
    b = ctx.se3_track_images_batch(refs, [im.ctypes.data for im in imgs], w, inits)
    for i in range(4):
        assert list(a[i].frameToRef) == list(b[i].frameToRef)
    ctx.close()


@pytest.mark.parametrize("mode", ["streamed", "per_chunk", "starved", "non_pinned_pitch"])
def test_host_image_pipeline_schedules(lsd, oracle, mode):
    """Every schedule of lsd_se3_track_images_batch with MORE THAN ONE chunk gives the poses of se3_track_batch bit for bit:
    the persistent tracker fed chunk by chunk (default), one tracker launch per chunk, the watchdog fallback (a streamed
    tracker whose producers never arrive stops itself and the batch is re-run: forced here with a 1 ns watchdog), and a
    pitched pageable source.  An argument error in the middle of the pipeline must return (not hang) and leave the context usable."""
    w, h = 320, 240
    n = 7
    ds = [make_oracle_pair(170 + s, w, h) for s in range(n)]
    ctx = lsd.Context(w, h, ds[0]["pr"]["K"])
    kfs = ctx.create_frames([d["kf_img"] for d in ds])
    for k, d in zip(kfs, ds):
        k.set_idepth(d["idepth"], d["var"])
    refs = ctx.create_refs(kfs)
    frs = ctx.create_frames([d["fr_img"] for d in ds])
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (n, 1))
    a = ctx.se3_track_batch(refs, frs, inits)
    pitch = w
    if mode == "non_pinned_pitch":
        pitch = w + 32
        imgs = []
        for d in ds:
            buf = np.full((h, pitch), 255, np.uint8)
            buf[:, :w] = d["fr_img"]
            imgs.append(buf)
        ctx.set_image_pipeline(chunk_frames=3, streamed=1)
    else:
        imgs = [np.ascontiguousarray(d["fr_img"]) for d in ds]
        if mode == "streamed":
            ctx.set_image_pipeline(chunk_frames=2, streamed=1)
        elif mode == "per_chunk":
            ctx.set_image_pipeline(chunk_frames=2, streamed=0)
        else:
            ctx.set_image_pipeline(chunk_frames=2, streamed=1, watchdog_seconds=1e-9)
    for rep in range(2):  # twice: the second call reuses pooled slabs and the tracker scratch
        b = ctx.se3_track_images_batch(refs, [im.ctypes.data for im in imgs], pitch, inits)
        for i in range(n):
            assert list(a[i].frameToRef) == list(b[i].frameToRef), (mode, rep, i)
            assert a[i].lastGoodCount == b[i].lastGoodCount and a[i].lastBadCount == b[i].lastBadCount
    # an argument error in the middle of the pipeline (a null image in the second chunk) must come back as an error, not hang,
    # and must leave the context usable
    ptrs = [im.ctypes.data for im in imgs]
    ptrs[3] = None
    with pytest.raises(lsd.LsdError):
        ctx.se3_track_images_batch(refs, ptrs, pitch, inits)
    b = ctx.se3_track_images_batch(refs, [im.ctypes.data for im in imgs], pitch, inits)
    assert all(list(a[i].frameToRef) == list(b[i].frameToRef) for i in range(n))
    ctx.close()


def test_permaref_quick_track_and_overlap_batch(lsd, oracle):
    """SE3Tracker::trackFrameOnPermaref / checkPermaRefOverlap: level-4 test track of n candidates
    in one launch.  referenceToFrame comes back un-inverted, LM traces match the oracle's evaluation by evaluation, and
    the tracked frames / keyframe counters are left untouched."""
    w, h = 320, 240
    ds = [make_oracle_pair(70 + s, w, h, max_t=0.05, max_r=np.radians(2.0)) for s in range(5)]
    ctx = lsd.Context(w, h, ds[0]["pr"]["K"])
    kfs = ctx.create_frames([d["kf_img"] for d in ds])
    frs = ctx.create_frames([d["fr_img"] for d in ds])
    for k, d in zip(kfs, ds):
        k.set_idepth(d["idepth"], d["var"])
    refs = ctx.create_refs(kfs)
    rng = np.random.default_rng(3)
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (5, 1))
    inits[:, 4:] = rng.normal(size=(5, 3)) * 0.01
    before = [(f.tracking_meta(), k.counters()) for f, k in zip(frs, kfs)]
    res, traces = ctx.se3_track_permaref_batch(refs, frs, inits, want_trace=True)
    usage = ctx.check_permaref_overlap_batch(refs, np.array([list(r.frameToRef) for r in res]))
    for i, d in enumerate(ds):
        ores, otrace = oracle.se3_track_permaref(d["oref"], d["ofr"], inits[i], 2)
        assert res[i].diverged == ores.diverged == 0
        assert res[i].trackingWasGood == ores.trackingWasGood
        assert [t[0] for t in traces[i]] == [4] * len(traces[i]) and len(traces[i]) == len(otrace) >= 2
        for tg, to in zip(traces[i], otrace):  # (level, accepted, error, lambda, bufSize)
            assert tg[1] == to[1] and tg[4] == to[4], (tg, to)
            assert abs(tg[2] - to[2]) <= RES_RTOL * abs(to[2]) + 1e-6
        pg, po = np.array(res[i].frameToRef), np.array(ores.frameToRef)
        assert np.linalg.norm(pg[4:] - po[4:]) <= POSE_TOL and quat_angle(pg[:4], po[:4]) <= POSE_TOL
        accepted = [t[2] for t in traces[i] if t[1] != 0]
        assert all(b < a for a, b in zip(accepted, accepted[1:])), "accepted LM steps must decrease the residual"
        ou = oracle.check_permaref_overlap(d["oref"], po)
        assert 0.3 < usage[i] <= 1.0 and abs(usage[i] - ou) <= 1e-5 * max(1.0, ou)
    after = [(f.tracking_meta(), k.counters()) for f, k in zip(frs, kfs)]
    for b, a in zip(before, after):
        assert b[0][0] == a[0][0] and np.array_equal(b[0][1], a[0][1]) and b[1] == a[1]
    # a batch of one equals its entry in the batch; the regular tracker is unaffected by the mode switch
    single = ctx.se3_track_permaref_batch([refs[2]], [frs[2]], [inits[2]])
    assert np.array_equal(np.array(single[0].frameToRef), np.array(res[2].frameToRef))
    full = ctx.se3_track(refs[2], frs[2], np.array([0, 0, 0, 1, 0, 0, 0.0]))
    assert full.trackingWasGood and sum(full.numResidualCalls[1:4]) > 0
    ctx.close()


def test_two_contexts_run_concurrently_on_one_device(lsd, oracle):
    """include/lsd_b200.h promises one context per calling thread with frame / reference handles shared read-only
    (upstream runs tracking, mapping and constraint search on separate threads: the reference's publishKeyframe takes a
    shared lock for exactly that reason, PangolinOutputIOWrapper.cpp:50).  Three host threads drive three contexts on
    the same device at the same time -- SE3 tracking, Sim3 tracking against the SAME references, and a depth-map update --
    and every result must equal the one obtained serially."""
    import threading
    from common import hyp_from_idepth, make_oracle_depth_scene, make_sim3_pair
    w, h = 320, 240
    n = 6
    ds = [make_sim3_pair(oracle, 300 + s, w, h) for s in range(n)]
    K = ds[0]["pr"]["K"]
    ctxA, ctxB, ctxC = lsd.Context(w, h, K), lsd.Context(w, h, K), lsd.Context(w, h, K)
    kfs = ctxA.create_frames([d["kf_img"] for d in ds])
    frs = ctxA.create_frames([d["fr_img"] for d in ds])
    frs_sim3 = ctxA.create_frames([d["fr_img"] for d in ds])  # Sim3 reads its frames only; SE3 writes the mask of its own
    for k, f, d in zip(kfs, frs_sim3, ds):
        k.set_idepth(d["idepth"], d["var"])
        f.set_idepth(d["fr_idepth"], d["fr_var"])
    refs = ctxA.create_refs(kfs)
    inits7 = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (n, 1))
    inits8 = np.array([d["gt8"] for d in ds])
    inits8[:, 7] = 1.0
    dsc = make_oracle_depth_scene(310, w, h, n_refs=4)
    kfC = ctxC.create_frame(dsc["kf_img"], 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    refC = []
    for i, r in enumerate(dsc["refs"]):
        f = ctxC.create_frame(r["img"], 1001 + i)
        f.set_tracking_meta(1000, r["toParent"], 1.0)
        refC.append(f)
    m0 = hyp_from_idepth(dsc["idepth"], dsc["var"])

    def se3():
        return [list(r.frameToRef) for r in ctxA.se3_track_batch(refs, frs, inits7)]

    def sim3():
        return [list(r.frameToRef) for r in ctxB.sim3_track_batch(refs, frs_sim3, inits8)]

    def depth():
        dm = ctxC.create_depthmap()
        dm.initializeFromMap(kfC, m0)
        kfC.set_depth_updated_flag(0)
        dm.updateKeyframe(refC)
        out = dm.read().tobytes()
        dm.destroy()
        return out

    want = (se3(), sim3(), depth())
    got = {0: [], 1: [], 2: []}
    errs = []

    def worker(k, fn, reps):
        try:
            for _ in range(reps):
                got[k].append(fn())
        except Exception as e:  # noqa: BLE001
            errs.append((k, repr(e)))

    ths = [threading.Thread(target=worker, args=(0, se3, 12)), threading.Thread(target=worker, args=(1, sim3, 12)),
           threading.Thread(target=worker, args=(2, depth, 6))]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=300)
    assert not errs, errs
    assert all(not t.is_alive() for t in ths)
    for k in range(3):
        assert got[k] and all(g == want[k] for g in got[k]), f"thread {k}: result changed under concurrency"
    for c in (ctxB, ctxC, ctxA):
        c.close()


@pytest.mark.parametrize("seed,wh", [(61, (640, 480)), (62, (1280, 960)), (63, (320, 240))])
def test_live_cluster_kernel_matches_queue_kernel(lsd, oracle, seed, wh):
    """k_se3_track_live (one thread-block cluster per pair: what a one-frame SlamSystem::trackFrame call runs on) against
    k_se3_track (the batch work-queue kernel) on the same pair: same records, same order inside a record, same record order,
    so every output -- pose, counters, the whole accept / reject trace, the refPixelWasGood mask -- must be identical bits,
    for every record size; and the live result stays within the pose tolerance of the EXACT oracle."""
    import ctypes as C
    w, h = wh
    d = make_oracle_pair(seed, w, h)
    ctx, kf, fr, ref = _gpu_pair(lsd, d, w, h)
    init = np.array([0, 0, 0, 1, 0, 0, 0.0])
    d2 = make_oracle_pair(seed + 100, w, h)
    kf2 = ctx.create_frame(d2["kf_img"], 2)
    fr2 = ctx.create_frame(d2["fr_img"], 3)
    kf2.set_idepth(d2["idepth"], d2["var"])
    ref2 = ctx.create_refs([kf2])[0]

    def snap(res, trace, frame):
        raw = bytes(C.string_at(C.addressof(res), C.sizeof(res)))
        rows = [tuple(t) for t in trace[: res.traceLen]]
        return raw, rows, frame.refPixelWasGood().copy()

    for pts in (0, 512, "live"):
        if pts == "live":  # per-level record sizes of a live context (lsd_ctx_set_live_tracking)
            ctx.set_live_tracking(True)
        else:
            ctx.set_se3_record_points(pts)
        ctx.set_se3_live_pairs(0)  # work-queue kernel
        rq, tq = ctx.se3_track(ref, fr, init, want_trace=True)
        q = snap(rq, tq, fr)
        rq2 = ctx.se3_track_batch([ref, ref2], [fr, fr2], [init, init])
        q2 = [bytes(C.string_at(C.addressof(r), C.sizeof(r))) for r in rq2]
        ctx.set_se3_live_pairs(2)  # cluster kernel for one and for two pairs
        rl, tl = ctx.se3_track(ref, fr, init, want_trace=True)
        l = snap(rl, tl, fr)
        assert l[1] == q[1], f"accept / reject trace differs at record size {pts}"
        assert l[0] == q[0], f"result struct differs at record size {pts}"
        assert np.array_equal(l[2], q[2]), "refPixelWasGood mask"
        rl2 = ctx.se3_track_batch([ref, ref2], [fr, fr2], [init, init])
        assert [bytes(C.string_at(C.addressof(r), C.sizeof(r))) for r in rl2] == q2, "two pairs = two clusters"
        ores, _ = oracle.se3_track(d["oref"], d["ofr"], init, 2)
        assert rl.diverged == ores.diverged and rl.trackingWasGood == ores.trackingWasGood
        assert np.linalg.norm(np.array(rl.frameToRef)[4:] - np.array(ores.frameToRef)[4:]) <= POSE_TOL
        assert quat_angle(np.array(rl.frameToRef)[:4], np.array(ores.frameToRef)[:4]) <= POSE_TOL
    ctx.set_live_tracking(False)
    ctx.set_se3_record_points(0)
    ctx.set_se3_live_pairs(-1)
    ctx.close()
