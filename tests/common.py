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
