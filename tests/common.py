"""Shared builders for the parity tests: the same seeded inputs go to the oracle and the CUDA path."""
import numpy as np

from lsd_b200 import synth
from oracle import pyoracle as O


def quat_angle(qa, qb):
    """rotation angle (rad) between two (x,y,z,w) quaternions"""
    qa = np.asarray(qa, np.float64)
    qb = np.asarray(qb, np.float64)
    if np.dot(qa, qb) < 0:
        qb = -qb
    # relative rotation qa^-1 * qb: vector part norm = sin(angle/2)
    w1, v1 = qa[3], -qa[:3]
    w2, v2 = qb[3], qb[:3]
    v = w1 * v2 + w2 * v1 + np.cross(v1, v2)
    return 2.0 * np.arcsin(min(1.0, float(np.linalg.norm(v))))


def make_oracle_pair(seed, w, h, var=0.01, noise=0.0, **kw):
    pr = synth.make_pair(seed, w, h, **kw)
    kf_img = pr["kf_img"].cpu().numpy()
    fr_img = pr["fr_img"].cpu().numpy()
    kf = O.Frame(2 * seed, kf_img, pr["K"])
    fr = O.Frame(2 * seed + 1, fr_img, pr["K"])
    kf.build_pyramids()
    fr.build_pyramids()
    mg = kf.get(O.MAXGRAD, 0)
    idv, vv = synth.semidense_idepth(pr["kf_depth"], mg, var=var, noise=noise, seed=seed)
    kf.set_idepth(idv, vv)
    ref = O.Ref(kf)
    return dict(pr=pr, kf_img=kf_img, fr_img=fr_img, okf=kf, ofr=fr, oref=ref, idepth=idv, var=vv)


def hyp_from_idepth(idepth, var, validity=20, smoothed=True):
    """A DepthMapPixelHypothesis map (32-byte AoS) from idepth / var planes (var <= 0: invalid pixel)."""
    h, w = idepth.shape
    m = np.zeros((h, w), O.HYP_DTYPE)
    valid = var > 0
    m["isValid"] = valid
    m["validity_counter"] = np.where(valid, validity, 0)
    m["idepth"] = np.where(valid, idepth, 0)
    m["idepth_var"] = np.where(valid, var, 0)
    m["idepth_smoothed"] = np.where(valid, idepth if smoothed else -1, 0)
    m["idepth_var_smoothed"] = np.where(valid, var if smoothed else -1, 0)
    return m


def make_oracle_depth_scene(seed, w, h, n_refs=10, noise=0.05, var=0.01, residual=1.0, with_mask=False, **kw):
    """Keyframe + reference frames for the DepthMap tests: oracle Frames with GT relative poses installed as
    thisToParent_raw (what SE3Tracker::trackFrame leaves behind), initialTrackedResidual = `residual`."""
    sc = synth.make_depth_scene(seed, w, h, n_refs, **kw)
    kf_img = sc["kf_img"].cpu().numpy()
    okf = O.Frame(1000, kf_img, sc["K"])
    okf.build_pyramids()
    mg = okf.get(O.MAXGRAD, 0)
    idv, vv = synth.semidense_idepth(sc["kf_depth"], mg, var=var, noise=noise, seed=seed)
    refs = []
    for i, r in enumerate(sc["refs"]):
        img = r["img"].cpu().numpy()
        f = O.Frame(1001 + i, img, sc["K"])
        f.build_pyramids()
        toParent = np.concatenate([r["toKf"], [1.0]])
        f.set_track_meta(residual, 1000, toParent)
        if with_mask:
            f.set_mask(np.ones((h >> 1, w >> 1), np.uint8))
        refs.append(dict(img=img, of=f, toParent=toParent, depth=r["depth"].cpu().numpy()))
    gt_idepth = (1.0 / sc["kf_depth"].cpu().numpy()).astype(np.float32)
    return dict(sc=sc, K=sc["K"], kf_img=kf_img, okf=okf, maxgrad=mg, idepth=idv, var=vv, refs=refs, gt_idepth=gt_idepth)


def make_sim3_pair(oracle, seed, w, h, c=1.0, var=0.01, max_t=0.05, max_r=np.radians(2.0), K=None):
    d = make_oracle_pair(seed, w, h, var=var, max_t=max_t, max_r=max_r, K=K)
    # frame B gets its own semi-dense depth, in a map whose inverse depths are c times the true ones
    mgB = d["ofr"].get(oracle.MAXGRAD, 0)
    idB, vB = synth.semidense_idepth(d["pr"]["fr_depth"], mgB, var=var)
    idB = np.where(vB > 0, idB * np.float32(c), idB).astype(np.float32)
    d["ofr"].set_idepth(idB, vB)
    gt = np.concatenate([d["pr"]["frameToRef"], [c]])  # p_ref = c * R * p_B(map) + t
    d["gt8"] = gt
    d["fr_idepth"], d["fr_var"] = idB, vB
    return d
