"""GPU parity for DepthMap (C1-C10) against the oracle, through the C ABI.

The kernels are compiled -fmad=false and restate every per-pixel statement in the oracle's operation order, so
the bar here is stricter than north_star's (<= 1e-3 relative on >= 99.5 % of valid pixels): validity masks,
blacklist counters, validity counters and skip-ahead ids must be bit-exact and so must every float of every
valid hypothesis.  The oracle runs with fp64 accumulation of its two whole-map sums (g_exactSums), the
order-independent definition the device implements as well.
"""
import ctypes

import numpy as np
import pytest

from common import hyp_from_idepth, make_oracle_depth_scene

pytestmark = pytest.mark.gpu

FLOATS = ("idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed", "nextStereoFrameMinID")


def assert_maps_equal(g, o, what):
    assert np.array_equal(g["isValid"], o["isValid"]), f"{what}: isValid differs at {np.argwhere(g['isValid'] != o['isValid'])[:5]}"
    assert np.array_equal(g["blacklisted"], o["blacklisted"]), f"{what}: blacklisted"
    v = o["isValid"] > 0
    assert np.array_equal(g["validity_counter"][v], o["validity_counter"][v]), f"{what}: validity_counter"
    for f in FLOATS:
        a, b = g[f][v], o[f][v]
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"{what}: {f} differs on {np.count_nonzero(~same)} of {v.sum()} valid pixels, max rel " \
                           f"{np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))}"


def gpu_scene(lsd, d, w, h):
    """Mirror an oracle depth scene on the device: keyframe + reference frames with the same bookkeeping."""
    ctx = lsd.Context(w, h, d["K"])
    kf = ctx.create_frame(d["kf_img"], 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
    refs = []
    for i, r in enumerate(d["refs"]):
        f = ctx.create_frame(r["img"], 1001 + i, flags=lsd.BUILD_MAXGRAD0)
        f.set_tracking_meta(1000, r["toParent"], 1.0)
        refs.append(f)
    return ctx, kf, refs


def _K_for(w, h):
    """BASELINE configs[4] runs at 1280x960 with the d2_camera.xml-style intrinsics."""
    from lsd_b200 import synth
    return synth.d2_K() if (w, h) == (1280, 960) else None


@pytest.mark.parametrize("wh,seed", [((320, 240), 31), ((640, 480), 32), ((640, 480), 37), ((1280, 960), 36)])
def test_stages_bit_exact(lsd, oracle, wh, seed):
    _stages_bit_exact(lsd, oracle, wh, seed, None)


@pytest.mark.parametrize("wh,seed,tma", [((176, 144), 41, 0), ((176, 144), 41, 3), ((320, 240), 31, 0), ((320, 240), 31, 3),
                                         ((640, 480), 32, 2)])
def test_stencil_tile_loads_tma_and_vector(lsd, oracle, wh, seed, tma):
    """lsd_ctx_set_stencil_tma: the TMA halo-tile path (40-wide boxes at a 16-byte aligned origin, hardware zero fill outside
    the map, partial tiles at the right / bottom edge) and the vector-load path produce the oracle's maps bit for bit."""
    _stages_bit_exact(lsd, oracle, wh, seed, tma)


def _stages_bit_exact(lsd, oracle, wh, seed, tma):
    w, h = wh
    oracle.set_exact_sums(1)
    d = make_oracle_depth_scene(seed, w, h, n_refs=10, K=_K_for(w, h))
    m0 = hyp_from_idepth(d["idepth"], d["var"])
    rng = np.random.default_rng(seed)
    holes = (rng.random((h, w)) < 0.15) & (m0["isValid"] > 0)
    m0["isValid"][holes] = 0
    m0["nextStereoFrameMinID"] = np.where(rng.random((h, w)) < 0.3, 1004.0, 0.0).astype(np.float32)  # exercise refById
    odm = oracle.DepthMap(w, h, d["K"])
    odm.init_map(d["okf"], m0)
    ctx, kf, refs = gpu_scene(lsd, d, w, h)
    if tma is not None:
        ctx.set_stencil_tma(tma)
    gdm = ctx.create_depthmap()
    gdm.initializeFromMap(kf, m0)
    assert_maps_equal(gdm.read(), odm.read(), "import/export round trip")

    odm.prepare([r["of"] for r in d["refs"]])
    gdm.prepare(refs)
    odm.stage(oracle.STAGE_OBSERVE)
    gdm.stage(lsd.STAGE_OBSERVE)
    go, oo = gdm.read(), odm.read()
    assert_maps_equal(go, oo, "observeDepth")
    assert (oo["idepth_var"] != m0["idepth_var"]).sum() > 0.1 * (m0["isValid"] > 0).sum(), "observe must have updated pixels"

    odm.stage(oracle.STAGE_FILL_HOLES)
    gdm.stage(lsd.STAGE_FILL_HOLES)
    assert_maps_equal(gdm.read(), odm.read(), "regularizeDepthMapFillHoles")
    odm.stage(oracle.STAGE_REGULARIZE, 0, 24)
    gdm.stage(lsd.STAGE_REGULARIZE, 0, 24)
    assert_maps_equal(gdm.read(), odm.read(), "regularizeDepthMap(false)")
    odm.stage(oracle.STAGE_REGULARIZE, 1, 24)
    gdm.stage(lsd.STAGE_REGULARIZE, 1, 24)
    assert_maps_equal(gdm.read(), odm.read(), "regularizeDepthMap(true)")

    odm.stage(oracle.STAGE_SET_DEPTH)
    gdm.stage(lsd.STAGE_SET_DEPTH)
    for l in range(5):
        assert np.array_equal(kf.idepth(l), d["okf"].get(oracle.IDEPTH, l)), f"idepth L{l}"
        assert np.array_equal(kf.idepthVar(l), d["okf"].get(oracle.IDEPTHVAR, l)), f"idepthVar L{l}"
    assert np.array_equal(gdm.debugPlotDepthMap(), odm.debug_rgb())

    # propagateDepth to the farthest frame, with and without the tracker's refPixelWasGood mask
    for with_mask in (False, True):
        d2 = make_oracle_depth_scene(seed, w, h, n_refs=10, K=_K_for(w, h))
        o2 = oracle.DepthMap(w, h, d2["K"])
        cur = odm.read()
        o2.init_map(d2["okf"], cur)
        g2 = ctx.create_depthmap()
        g2.initializeFromMap(kf, cur)
        new_o, new_g = d2["refs"][-1]["of"], refs[-1]
        if with_mask:
            mask = (np.random.default_rng(5).random((h >> 1, w >> 1)) < 0.9).astype(np.uint8)
            new_o.set_mask(mask)
            new_g.set_mask(mask)
        else:
            new_g.set_mask(None)
        o2.stage(oracle.STAGE_PROPAGATE, frame=new_o)
        g2.stage(lsd.STAGE_PROPAGATE, frame=new_g)
        assert_maps_equal(g2.read(), o2.read(), f"propagateDepth(mask={with_mask})")
        g2.destroy()
    oracle.set_exact_sums(0)
    ctx.close()


@pytest.mark.parametrize("back,expect_max", [(2.5, 5), (6.0, 9)])
def test_propagate_many_sources_per_target(lsd, oracle, back, expect_max):
    """propagateDepth onto a view far BEHIND the keyframe: the map shrinks towards the image centre and up to a dozen sources
    land on one target pixel, so the replay's three paths all run -- one record, 2..4 records ordered in registers, and the
    overflow list of a fifth and later arrival -- and upstream's raster-order merge / occlusion rules must come out bit for bit."""
    w, h = 320, 240
    oracle.set_exact_sums(1)
    d = make_oracle_depth_scene(52, w, h, n_refs=2, noise=0.0)
    m0 = hyp_from_idepth(d["idepth"], d["var"])
    odm = oracle.DepthMap(w, h, d["K"])
    odm.init_map(d["okf"], m0)
    ctx, kf, refs = gpu_scene(lsd, d, w, h)
    gdm = ctx.create_depthmap()
    gdm.initializeFromMap(kf, m0)
    # the new frame sits `back` metres behind the keyframe, same orientation: thisToParent = (identity, (0, 0, -back))
    toParent = np.array([0, 0, 0, 1, 0, 0, -back, 1.0])
    new_o, new_g = d["refs"][-1]["of"], refs[-1]
    new_o.set_track_meta(1.0, 1000, toParent)
    new_g.set_tracking_meta(1000, toParent, 1.0)
    mask = np.ones((h >> 1, w >> 1), np.uint8)  # tracked-on-this-keyframe path: no photometric test between the two views
    new_o.set_mask(mask)
    new_g.set_mask(mask)
    # how many sources per target this really is (same projection, float64): the test must reach the overflow path
    fx, fy, cx, cy = d["K"][:4]
    ys, xs = np.nonzero(m0["isValid"] > 0)
    z = 1.0 / m0["idepth_smoothed"][ys, xs].astype(np.float64)
    u = (xs - cx) / fx * z / (z + back) * fx + cx
    v = (ys - cy) / fy * z / (z + back) * fy + cy
    cnt = np.bincount(((u + 0.5).astype(int) + (v + 0.5).astype(int) * w), minlength=w * h)
    assert cnt.max() >= expect_max, cnt.max()
    odm.stage(oracle.STAGE_PROPAGATE, frame=new_o)
    gdm.stage(lsd.STAGE_PROPAGATE, frame=new_g)
    go, oo = gdm.read(), odm.read()
    assert_maps_equal(go, oo, f"propagateDepth, {back} m behind (up to {cnt.max()} sources per target)")
    assert (oo["isValid"] > 0).sum() > 200
    # and the scratch is clean again: a second propagate of the same map gives the same result
    g2 = ctx.create_depthmap()
    g2.initializeFromMap(kf, m0)
    g2.stage(lsd.STAGE_PROPAGATE, frame=new_g)
    gdm.initializeFromMap(kf, m0)
    gdm.stage(lsd.STAGE_PROPAGATE, frame=new_g)
    assert_maps_equal(gdm.read(), g2.read(), "second propagate on the same scratch")
    oracle.set_exact_sums(0)
    ctx.close()


@pytest.mark.parametrize("wh", [(320, 240), (640, 480), (1280, 960)])
def test_update_and_create_keyframe_sequence(lsd, oracle, wh):
    """The live mapping loop: 8 x updateKeyframe (one tracked frame each), then createKeyFrame on the 9th, then 1 update.
    Run at every BASELINE resolution (640x480: configs[0..3]; 1280x960 with d2 intrinsics: configs[4])."""
    w, h = wh
    oracle.set_exact_sums(1)
    d = make_oracle_depth_scene(33, w, h, n_refs=10, with_mask=True, K=_K_for(w, h))
    ctx, kf, refs = gpu_scene(lsd, d, w, h)
    for f in refs:
        f.set_mask(np.ones((h >> 1, w >> 1), np.uint8))
    d["okf"].set_depth_gt(d["sc"]["kf_depth"].cpu().numpy())
    kf.set_depth_from_gt(d["sc"]["kf_depth"].cpu().numpy())
    odm, gdm = oracle.DepthMap(w, h, d["K"]), ctx.create_depthmap()
    odm.init_gt(d["okf"])
    gdm.initializeFromGTDepth(kf)
    assert_maps_equal(gdm.read(), odm.read(), "initializeFromGTDepth")
    d["okf"].set_counters(3, 0)
    kf.set_counters(3, 0)
    for i in range(8):
        odm.update_keyframe([d["refs"][i]["of"]])
        gdm.updateKeyframe([refs[i]])
        assert_maps_equal(gdm.read(), odm.read(), f"updateKeyframe #{i}")
    assert kf.counters() == (3, 8) and tuple(oracle.frame_counters(d["okf"])[:2]) == (3, 8)
    # two frames in one call (deque of 2) exercises referenceFrameByID
    odm.update_keyframe([d["refs"][7]["of"], d["refs"][8]["of"]])
    gdm.updateKeyframe([refs[7], refs[8]])
    assert_maps_equal(gdm.read(), odm.read(), "updateKeyframe(deque of 2)")

    new_o, new_g = d["refs"][9]["of"], refs[9]
    odm.create_keyframe(new_o)
    f = gdm.createKeyFrame(new_g)
    assert f == odm.last_rescale(), (f, odm.last_rescale())
    assert_maps_equal(gdm.read(), odm.read(), "createKeyFrame")
    pid, pose, _ = new_g.tracking_meta()
    assert np.allclose(pose, oracle.frame_pose(new_o), rtol=0, atol=1e-12)
    for l in range(5):
        assert np.array_equal(new_g.idepth(l), new_o.get(oracle.IDEPTH, l)), f"new keyframe idepth L{l}"
        assert np.array_equal(new_g.idepthVar(l), new_o.get(oracle.IDEPTHVAR, l))
    mi, npts = new_g.mean_idepth()
    assert npts == new_o.num_points() and abs(mi - new_o.mean_idepth()) <= 1e-6 * abs(mi)
    odm.finalize()
    gdm.finalizeKeyFrame()
    assert_maps_equal(gdm.read(), odm.read(), "finalizeKeyFrame")
    oracle.set_exact_sums(0)
    ctx.close()


def test_create_from_empty_map_and_blacklisting(lsd, oracle):
    """observeDepthCreate over the full idepth range on an empty map; repeated failures walk the blacklist counter."""
    w, h = 320, 240
    d = make_oracle_depth_scene(34, w, h, n_refs=6, step=0.02)
    ctx, kf, refs = gpu_scene(lsd, d, w, h)
    empty = hyp_from_idepth(np.zeros((h, w), np.float32), -np.ones((h, w), np.float32))
    odm, gdm = oracle.DepthMap(w, h, d["K"]), ctx.create_depthmap()
    odm.init_map(d["okf"], empty)
    gdm.initializeFromMap(kf, empty)
    for i in (5, 3, 1):
        odm.prepare([d["refs"][i]["of"]])
        gdm.prepare([refs[i]])
        odm.stage(oracle.STAGE_OBSERVE)
        gdm.stage(lsd.STAGE_OBSERVE)
        assert_maps_equal(gdm.read(), odm.read(), f"create pass with frame {i}")
    m = gdm.read()
    assert (m["isValid"] > 0).sum() > 0.05 * w * h and (m["blacklisted"] < 0).any()
    ctx.close()


def test_initialize_randomly_consumes_rand_like_upstream(lsd, oracle):
    w, h = 320, 240
    d = make_oracle_depth_scene(35, w, h, n_refs=1)
    ctx, kf, refs = gpu_scene(lsd, d, w, h)
    odm, gdm = oracle.DepthMap(w, h, d["K"]), ctx.create_depthmap()
    odm.init_random(d["okf"], seed=7)
    ctypes.CDLL(None).srand(7)
    gdm.initializeRandomly(kf)
    assert_maps_equal(gdm.read(), odm.read(), "initializeRandomly")
    assert np.array_equal(kf.idepth(0), d["okf"].get(oracle.IDEPTH, 0))
    ctx.close()


def test_batched_stage_equals_single(lsd, oracle):
    """blockIdx.z batching: n maps through one set of launches == the same maps one at a time."""
    w, h = 320, 240
    ds = [make_oracle_depth_scene(40 + i, w, h, n_refs=3) for i in range(3)]
    ctx = lsd.Context(w, h, ds[0]["K"])
    maps, singles = [], []
    for d in ds:
        kf = ctx.create_frame(d["kf_img"], 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
        refs = []
        for i, r in enumerate(d["refs"]):
            f = ctx.create_frame(r["img"], 1001 + i)
            f.set_tracking_meta(1000, r["toParent"], 1.0)
            refs.append(f)
        m0 = hyp_from_idepth(d["idepth"], d["var"])
        for lst in (maps, singles):
            dm = ctx.create_depthmap()
            dm.initializeFromMap(kf, m0)
            dm.prepare(refs)
            lst.append(dm)
    for stage, a1, a2 in [(lsd.STAGE_OBSERVE, 0, 0), (lsd.STAGE_FILL_HOLES, 0, 0), (lsd.STAGE_REGULARIZE, 0, 24)]:
        ctx.depth_stage_batch(maps, stage, a1, a2)
        for dm in singles:
            dm.stage(stage, a1, a2)
    for a, b in zip(maps, singles):
        assert a.read().tobytes() == b.read().tobytes()
    ctx.close()
