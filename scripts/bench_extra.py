#!/usr/bin/env python
"""Secondary measurements (BASELINE.json configs[2] and configs[3]); bench.py embeds the result under "extra".

  depthmap: per-stage device time of DepthMap on B independent 640x480 keyframes with 10 reference frames each
            (one set of launches, blockIdx.z = keyframe; B x ~27 MB working set > 126 MB L2), algorithmic bytes per
            SURVEY.md 8(d) config 3, the latency of one updateKeyframe / createKeyFrame call, and the oracle port
            on 4 host threads (upstream MAPPING_THREADS) beside it.
  sim3:     constraint-search shaped batch: 64 candidates x 2 directions of trackFrameSim3 (levels 4 -> 1) in one launch,
            vs the oracle port on all host threads.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

W, H = 640, 480
N = W * H
PEAK_FALLBACK = 6650.0


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return PEAK_FALLBACK, "fallback"


def depth_bench(ctx, lsd, B=64, n_refs=10, reps=5, device="cuda", cpu=True):
    import torch

    from lsd_b200 import synth
    from common import hyp_from_idepth
    K = synth.default_K(W, H)
    scenes = [synth.make_depth_scene(500 + i, W, H, n_refs, K=K, device=device) for i in range(B)]
    kfs, refs, maps0, dms, new_kfs = [], [], [], [], []
    rng = np.random.default_rng(0)
    for i, sc in enumerate(scenes):
        kf = ctx.create_frame(sc["kf_img"].cpu().numpy(), 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
        mg = kf.maxGradients(0)
        idv, vv = synth.semidense_idepth(sc["kf_depth"], mg, var=0.01, noise=0.05, seed=i)
        rr = []
        for j, r in enumerate(sc["refs"]):
            f = ctx.create_frame(r["img"].cpu().numpy(), 1001 + j, flags=lsd.BUILD_MAXGRAD0)
            f.set_tracking_meta(1000, np.concatenate([r["toKf"], [1.0]]), 1.0)
            rr.append(f)
        m0 = hyp_from_idepth(idv, vv)
        dm = ctx.create_depthmap()
        kfs.append(kf); refs.append(rr); maps0.append(m0); dms.append(dm)
    valid_frac = float(np.mean([m["isValid"].mean() for m in maps0]))

    def reset():
        for dm, kf, m0, rr in zip(dms, kfs, maps0, refs):
            dm.initializeFromMap(kf, m0)
            dm.prepare(rr)

    def timed(stage, a1=0, a2=0, frames=None, pre=None):
        ts = []
        for _ in range(reps):
            reset()
            if pre:
                pre()
            ctx.depth_stage_batch(dms, stage, a1, a2, frames)
            ts.append(ctx.last_stage_ms())
        return float(np.median(ts))

    out = {"B": B, "n_refs": n_refs, "valid_fraction": valid_frac}
    pk, src = peak()
    # success fraction of stereo for the byte model: pixels whose variance changed
    reset()
    ctx.depth_stage_batch(dms, lsd.STAGE_OBSERVE)
    after = dms[0].read()
    s = float(((after["idepth_var"] != maps0[0]["idepth_var"]) & (maps0[0]["isValid"] > 0)).mean())
    bpp = {
        "observeDepth": 29 + 29 + 4 + 4 + 4 * 1 + 16 * s + 0.25,   # R_used = 1 (nextStereoFrameMinID = 0 -> oldest frame)
        "fillHoles": 46.0,
        "regularize(false)": 30.0,
        "regularize(true)": 30.0,
        "propagateDepth": 29 * valid_frac + 12 + 29 * valid_frac + 5,
        "setDepth+pyramids": 9 + 8 * (409200 / 307200),
    }
    stages = {}
    stages["observeDepth"] = timed(lsd.STAGE_OBSERVE)
    stages["fillHoles"] = timed(lsd.STAGE_FILL_HOLES)
    stages["regularize(false)"] = timed(lsd.STAGE_REGULARIZE, 0, 24)
    stages["regularize(true)"] = timed(lsd.STAGE_REGULARIZE, 1, 24)
    stages["setDepth+pyramids"] = timed(lsd.STAGE_SET_DEPTH)
    new_frames = [rr[-1] for rr in refs]
    stages["propagateDepth"] = timed(lsd.STAGE_PROPAGATE, frames=new_frames)
    out["stages"] = {k: {"ms_batch": v, "us_per_keyframe": 1e3 * v / B, "alg_bytes_per_px": bpp[k],
                         "achieved_GBs": bpp[k] * N * B / (v * 1e-3) / 1e9, "frac_of_" + src + "_peak": bpp[k] * N * B / (v * 1e-3) / 1e9 / pk}
                     for k, v in stages.items()}
    # latency of the two upstream calls on ONE keyframe (blocking API, includes launch + sync overheads)
    lat_u, lat_c = [], []
    for _ in range(reps):
        reset()
        kfs[0].set_depth_updated_flag(0)
        t0 = time.perf_counter(); dms[0].updateKeyframe(refs[0]); lat_u.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); dms[0].createKeyFrame(refs[0][-1]); lat_c.append(time.perf_counter() - t0)
    out["latency_ms"] = {"updateKeyframe": 1e3 * float(np.median(lat_u)), "createKeyFrame": 1e3 * float(np.median(lat_c))}
    upd_batch = stages["observeDepth"] + stages["fillHoles"] + stages["regularize(false)"] + stages["setDepth+pyramids"]
    out["updateKeyframe_batched_keyframes_per_s"] = B / (upd_batch * 1e-3)

    if cpu:
        from oracle import pyoracle as O
        O.build()
        sc = scenes[0]
        okf = O.Frame(1000, sc["kf_img"].cpu().numpy(), K, fast=True)
        okf.build_pyramids()
        ofr = []
        for j, r in enumerate(sc["refs"]):
            f = O.Frame(1001 + j, r["img"].cpu().numpy(), K, fast=True)
            f.build_pyramids()
            f.set_track_meta(1.0, 1000, np.concatenate([r["toKf"], [1.0]]))
            ofr.append(f)
        odm = O.DepthMap(W, H, K, threads=4, fast=True)
        cpu_ms = {}
        def ctimed(stage, a1=0, a2=0, frame=None):
            ts = []
            for _ in range(3):
                odm.init_map(okf, maps0[0])
                odm.prepare(ofr)
                ts.append(odm.stage(stage, a1, a2, frame))
            return 1e3 * float(np.median(ts))
        cpu_ms["observeDepth"] = ctimed(O.STAGE_OBSERVE)
        cpu_ms["fillHoles"] = ctimed(O.STAGE_FILL_HOLES)
        cpu_ms["regularize(false)"] = ctimed(O.STAGE_REGULARIZE, 0, 24)
        cpu_ms["regularize(true)"] = ctimed(O.STAGE_REGULARIZE, 1, 24)
        cpu_ms["setDepth+pyramids"] = ctimed(O.STAGE_SET_DEPTH)
        cpu_ms["propagateDepth"] = ctimed(O.STAGE_PROPAGATE, frame=ofr[-1])
        out["cpu_port_ms_per_keyframe"] = {"threads": 4, "kind": "port", **cpu_ms}
    for dm in dms:
        dm.destroy()
    return out


def sim3_bench(ctx, lsd, n_cand=64, reps=3, device="cuda", cpu=True):
    from lsd_b200 import synth
    K = synth.default_K(W, H)
    kf_imgs, fr_imgs, idA, vA, idB, vB, gts = [], [], [], [], [], [], []
    import torch
    for i in range(n_cand):
        pr = synth.make_pair(900 + i, W, H, K, device=device, max_t=0.10, max_r=np.radians(3.0))
        kf_imgs.append(pr["kf_img"].cpu().numpy()); fr_imgs.append(pr["fr_img"].cpu().numpy())
        gts.append(np.concatenate([pr["frameToRef"], [1.0]]))
        kfd, frd = pr["kf_depth"], pr["fr_depth"]
        idA.append(kfd); idB.append(frd)
    A = ctx.create_frames(kf_imgs, flags=lsd.BUILD_MAXGRAD0)
    Bf = ctx.create_frames(fr_imgs, flags=lsd.BUILD_MAXGRAD0)
    host = []
    for i in range(n_cand):
        a_id, a_v = synth.semidense_idepth(idA[i], A[i].maxGradients(0))
        b_id, b_v = synth.semidense_idepth(idB[i], Bf[i].maxGradients(0))
        A[i].set_idepth(a_id, a_v); Bf[i].set_idepth(b_id, b_v)
        host.append((a_id, a_v, b_id, b_v))
    refA, refB = ctx.create_refs(A), ctx.create_refs(Bf)
    from lsd_b200.pipeline import sim3_inv
    rng = np.random.default_rng(1)
    inits_ab = np.array(gts)
    inits_ab[:, 4:7] += rng.normal(size=(n_cand, 3)) * 0.01
    inits_ba = np.array([sim3_inv(g) for g in inits_ab])
    refs = refA + refB
    frames = Bf + A
    inits = np.concatenate([inits_ab, inits_ba])
    ts, kms = [], []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        res = ctx.sim3_track_batch(refs, frames, inits, 4, 1)
        ts.append(time.perf_counter() - t0)
        kms.append(ctx.se3_last_stats())
    t = float(np.median(ts[1:]))
    byts, evals, kms_ = kms[-1]
    pk, src = peak()
    scale_err = float(np.median([abs(res[i].frameToRef[7] - 1.0) for i in range(n_cand)]))
    terr = float(np.median([np.linalg.norm(np.array(res[i].frameToRef[4:7]) - gts[i][4:7]) for i in range(n_cand)]))
    out = {"candidates": n_cand, "tracks_per_launch": 2 * n_cand, "levels": "4->1", "ms_per_search": 1e3 * t,
           "candidates_per_s": n_cand / t, "kernel_ms": kms_, "evaluations": evals, "alg_bytes": byts,
           "achieved_GBs": byts / (kms_ * 1e-3) / 1e9, "frac_of_" + src + "_peak": byts / (kms_ * 1e-3) / 1e9 / pk,
           "diverged": int(sum(r.diverged for r in res)), "median_scale_err": scale_err, "median_translation_err_m": terr}
    # the cheap pre-filter of the same search: SE3Tracker::trackFrameOnPermaref (level 4 only) + checkPermaRefOverlap
    pinits = np.concatenate([inits_ba[:, :7], inits_ab[:, :7]])
    tp = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        pres = ctx.se3_track_permaref_batch(refs, frames, pinits)
        usage = ctx.check_permaref_overlap_batch(refs, np.array([list(r.frameToRef) for r in pres]))
        tp.append(time.perf_counter() - t0)
    out["permaref_quick_check"] = {"tracks": 2 * n_cand, "ms": 1e3 * float(np.median(tp[1:])), "kernel_ms": ctx.se3_last_stats()[2],
                                   "tracks_per_s": 2 * n_cand / float(np.median(tp[1:])), "good": int(sum(r.trackingWasGood for r in pres)),
                                   "mean_overlap": float(usage.mean())}
    if cpu:
        from oracle import pyoracle as O
        O.build()
        threads = os.cpu_count() or 1
        m = min(n_cand, 16)
        oA, oB = [], []
        for i in range(m):
            a = O.Frame(2 * i, kf_imgs[i], K, fast=True); b = O.Frame(2 * i + 1, fr_imgs[i], K, fast=True)
            a.build_pyramids(); b.build_pyramids()
            a.set_idepth(host[i][0], host[i][1]); b.set_idepth(host[i][2], host[i][3])
            oA.append(a); oB.append(b)
        orA, orB = [O.Ref(a) for a in oA], [O.Ref(b) for b in oB]
        for r in orA + orB:
            for l in (1, 2, 3, 4):
                r.num(l)
        ii = np.concatenate([inits_ab[:m], inits_ba[:m]])
        secs, outs = O.sim3_track_batch(orA + orB, oB + oA, ii, 4, 1, 0, threads)
        t0 = time.perf_counter()
        for i in range(2 * m):
            O.se3_track_permaref((orA + orB)[i], (oB + oA)[i], np.concatenate([inits_ba[:m, :7], inits_ab[:m, :7]])[i], 0)
        out["permaref_quick_check"]["cpu_port_tracks_per_s"] = 2 * m / (time.perf_counter() - t0)
        out["cpu_port"] = {"candidates_per_s": m / secs, "threads": threads, "kind": "port", "sample": f"{m} candidates x 2 directions",
                           "max_scale_diff_vs_gpu": float(max(abs(outs[i].frameToRef[7] - res[i].frameToRef[7]) for i in range(m)))}
    return out


def vbo_bench(ctx, lsd, B=64, reps=5, device="cuda", cpu=True):
    """Keyframe::computeVbo on B keyframes in one launch; algorithmic bytes = 12 B/px read (idepth, var,
    image) + 16 B per emitted vertex.  CPU leg: the reference's OWN Keyframe.h (oracle/_ref) when present, else the port."""
    from lsd_b200 import synth
    K = synth.default_K(W, H)
    frames, host = [], []
    for i in range(B):
        pr = synth.make_pair(700 + i, W, H, K, device=device)
        f = ctx.create_frame(pr["kf_img"].cpu().numpy(), i, flags=lsd.BUILD_MAXGRAD0)
        idv, vv = synth.semidense_idepth(pr["kf_depth"], f.maxGradients(0), var=1e-4, noise=0.002, seed=i)
        idv = np.where(vv > 0, idv, -1).astype(np.float32)  # what Frame::setDepth leaves on invalid pixels
        vv = np.where(vv > 0, vv, -1).astype(np.float32)
        f.set_idepth(idv, vv)
        frames.append(f)
        if i == 0:
            host = (idv, vv, pr["kf_img"].cpu().numpy().astype(np.float32))
    scales = np.ones(B, np.float32)
    ts = []
    for _ in range(reps + 1):
        pts, _ = ctx.compute_vbo_batch(frames, scales, read=False)
        ts.append(ctx.last_stage_ms())
    ms = float(np.median(ts[1:]))
    byts = 12.0 * N * B + 16.0 * float(pts.sum())
    pk, src = peak()
    out = {"B": B, "vertices_per_keyframe": float(pts.mean()), "ms_batch": ms, "us_per_keyframe": 1e3 * ms / B, "alg_bytes": byts,
           "achieved_GBs": byts / (ms * 1e-3) / 1e9, "frac_of_" + src + "_peak": byts / (ms * 1e-3) / 1e9 / pk}
    t0 = time.perf_counter(); one = frames[0].compute_vbo(1.0); out["latency_ms_one_keyframe_incl_d2h"] = 1e3 * (time.perf_counter() - t0)
    if cpu:
        from oracle import pyoracle as O
        O.build()
        pk_pts = O.publish_keyframe_pack(*host)
        Kf = np.array(K, np.float32)
        kind = "reference" if O.ref_keyframe_lib() is not None else "port"
        fn = (lambda: O.ref_compute_vbo(pk_pts, Kf, 1.0)) if kind == "reference" else (lambda: O.compute_vbo(pk_pts, Kf, 1.0, fast=True))
        fn()
        t0 = time.perf_counter()
        for _ in range(10):
            v = fn()
        out["cpu"] = {"ms_per_keyframe": 1e2 * (time.perf_counter() - t0), "threads": 1, "kind": kind,
                      "bit_exact_vs_gpu": bool(v.tobytes() == one.tobytes())}
    for f in frames:
        f.release()
    return out


def pipeline_bench(lsd, w=640, h=480, n_frames=500, cpu_frames=60, device="cuda", cpu=True, K=None, contrast=60.0):
    """BASELINE configs[0] / configs[4] shape: lock-step tracking + mapping (lsd_b200/pipeline.py: track every frame, map
    every frame, keyframe switch by the upstream score) over a synthetic rendered sequence with known trajectory, through
    the blocking C ABI with HOST images (H2D inside the timed loop).  The CPU leg runs the SAME driver on the oracle port
    (1 tracking thread + 4 mapping threads = upstream MAPPING_THREADS) over the first `cpu_frames` frames."""
    from lsd_b200 import synth
    from lsd_b200.pipeline import DeviceBackend, LockStepSlam
    K = K or synth.default_K(w, h)
    room = synth.make_room(0, device=device, contrast=contrast)
    traj = synth.trajectory(n_frames, seed=0)
    frames = []
    for i, (R, t) in enumerate(traj):
        img, depth = synth.render(room, w, h, K, R, t, noise_seed=i)
        frames.append((img.cpu().numpy(), depth.cpu().numpy() if i == 0 else None))

    def run(backend, n):
        slam = LockStepSlam(backend)
        slam.first_frame(frames[0][0], 0, frames[0][1])
        t_kf, n_kf0 = 0.0, 0
        t0 = time.perf_counter()
        for i in range(1, n):
            k0 = slam.stats["keyframes"]
            t1 = time.perf_counter()
            slam.next_image(frames[i][0], i)
            if slam.stats["keyframes"] != k0:
                t_kf += time.perf_counter() - t1
        dt = time.perf_counter() - t0
        return slam, dt, t_kf

    class Native:  # csrc/slam.cu through the LockStepSlam-shaped surface run() needs
        def __init__(self, ctx):
            self.s = lsd.SlamSystem(ctx, keep_keyframes=False)
            self.stats = dict(tracked=0, lost=0, keyframes=0)
            self.world_poses = []

        def first_frame(self, img, fid, depth):
            st = self.s.gtDepthInit(img, fid, depth)
            self.world_poses.append((fid, np.array(st.camToWorld)))

        def next_image(self, img, fid):
            st = self.s.nextImage(img, fid)
            if st.tracked:
                self.world_poses.append((fid, np.array(st.camToWorld)))
                self.stats["tracked"] += 1
                self.stats["keyframes"] += st.isKeyframe
            else:
                self.stats["lost"] += 1

    def run_native(ctx, n):
        slam = Native(ctx)
        slam.first_frame(frames[0][0], 0, frames[0][1])
        t_kf = 0.0
        t0 = time.perf_counter()
        for i in range(1, n):
            k0 = slam.stats["keyframes"]
            t1 = time.perf_counter()
            slam.next_image(frames[i][0], i)
            if slam.stats["keyframes"] != k0:
                t_kf += time.perf_counter() - t1
        dt = time.perf_counter() - t0
        slam.stage_seconds = slam.s.stage_seconds()
        slam.s.close()
        return slam, dt, t_kf

    ctx = lsd.Context(w, h, K, device=0)
    rec = int(os.environ.get("EXTRA_SE3_RECORD", "1024"))
    ctx.set_se3_record_points(rec)  # a context that tracks one live sequence: small records (include/lsd_b200.h)
    run_native(ctx, min(20, n_frames))  # warm-up (allocations, pools, lazy init)
    slam, dt, t_kf = run_native(ctx, n_frames)
    pslam, pdt, _ = run(DeviceBackend(ctx), min(n_frames, 100))  # the Python driver over the same ABI, for reference
    R0, t0_ = traj[0]
    gt = np.array([R0.T @ (t - t0_) for _, t in traj])
    ids = [i for i, _ in slam.world_poses]
    est = np.array([p[4:7] for _, p in slam.world_poses])
    ate = float(np.sqrt(np.mean(np.sum((est - gt[ids]) ** 2, axis=1))))
    out = {"width": w, "height": h, "texture_contrast": contrast, "frames": n_frames, "fps": (n_frames - 1) / dt, "ms_per_frame": 1e3 * dt / (n_frames - 1),
           "keyframes": slam.stats["keyframes"], "lost": slam.stats["lost"],
           "ms_per_keyframe_switch_frame": 1e3 * t_kf / max(1, slam.stats["keyframes"]),
           "ate_rmse_m": ate, "path_m": float(np.linalg.norm(np.diff(gt, axis=0), axis=1).sum()),
           "h2d_bytes_per_frame": w * h, "se3_record_points": rec, "driver": "native lock-step driver (csrc/slam.cu: lsd_slam_next_image), blocking, host images",
           "python_driver_fps": (min(n_frames, 100) - 1) / pdt,
           "stage_ms_per_frame": {k: 1e3 * v / (n_frames - 1) for k, v in slam.stage_seconds.items()}}
    ctx.close()
    if cpu:
        from oracle import pyoracle as O
        from oracle_backend import OracleBackend
        O.build()
        m = min(cpu_frames, n_frames)
        oslam, odt, okf = run(OracleBackend(w, h, K, mode=0, threads=4, fast=True), m)
        both = min(len(oslam.world_poses), len(slam.world_poses), m)
        dev = float(np.abs(np.array([p[4:7] for _, p in oslam.world_poses[:both]]) - est[:both]).max())
        out["cpu_port"] = {"fps": (m - 1) / odt, "ms_per_frame": 1e3 * odt / (m - 1), "frames": m, "kind": "port",
                           "threads": "1 tracking + 4 mapping (upstream MAPPING_THREADS)", "keyframes": oslam.stats["keyframes"],
                           "max_translation_diff_vs_gpu_m": dev}
        out["speedup_vs_cpu_port"] = out["fps"] / out["cpu_port"]["fps"]
    return out


def main():
    import torch

    import lsd_b200
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from lsd_b200 import synth
    ctx = lsd_b200.Context(W, H, synth.default_K(W, H), device=0)
    reps = int(os.environ.get("EXTRA_REPS", "5"))
    cpu = os.environ.get("EXTRA_NO_CPU", "") == ""
    parts = os.environ.get("EXTRA_PARTS", "depthmap,sim3,vbo,pipeline").split(",")
    B = int(os.environ.get("EXTRA_B", "64"))
    out = {}
    if "depthmap" in parts:
        out["depthmap"] = depth_bench(ctx, lsd_b200, B=B, reps=reps, cpu=cpu)
    if "sim3" in parts:
        out["sim3"] = sim3_bench(ctx, lsd_b200, reps=min(reps, 3), cpu=cpu)
    if "vbo" in parts:
        out["vbo"] = vbo_bench(ctx, lsd_b200, B=B, reps=reps, cpu=cpu)
    from lsd_b200 import synth
    nf = int(os.environ.get("EXTRA_FRAMES", "500"))
    if "pipeline" in parts:
        out["pipeline_640x480"] = pipeline_bench(lsd_b200, 640, 480, nf, cpu=cpu)
    if "pipeline_d2" in parts:
        out["pipeline_1280x960"] = pipeline_bench(lsd_b200, 1280, 960, nf, cpu_frames=20, cpu=cpu, K=synth.d2_K(),
                                                   contrast=130.0)  # same wall texture seen at twice the resolution: per-pixel gradients halve, so the contrast is raised to keep ~the same semi-dense density
    ctx.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
