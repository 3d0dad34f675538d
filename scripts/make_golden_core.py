#!/usr/bin/env python
"""Regression fixtures of the UNPINNED core restatement (oracle/: Frame pyramids, SE3 / Sim3 trackers, DepthMap).

The reference holds no golden vector for this path and its implementation cannot be built here, so
these vectors do NOT pin parity with the reference: they freeze the restatement's own outputs on small seeded inputs so
that (a) the oracle cannot drift silently between rounds and (b) the CUDA path is checked against committed numbers as well
as against the live oracle.  Integer / mask / count planes are stored exactly (or as SHA-256), floating point as values.
Run:  python scripts/make_golden_core.py     (writes tests/golden/core_regression.npz)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"), os.path.join(ROOT, "tests")]
from common import hyp_from_idepth, make_oracle_depth_scene, make_oracle_pair, make_sim3_pair  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


PAIR_W, PAIR_H = 320, 240


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def canonical_map(m):
    """Hypothesis map with the fields upstream never reads on invalid pixels zeroed (they differ between implementations)."""
    c = m.copy()
    inv = c["isValid"] == 0
    for f in ("validity_counter", "idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed", "nextStereoFrameMinID"):
        c[f][inv] = 0
    c["_pad"] = 0
    return c


def build():
    out = {}
    w, h = 160, 112
    # ---- Frame pyramids + SE3 tracker (EXACT accumulation mode = the order-independent value).  320x240: level 4 of a
    # smaller image holds a few dozen points and the LM path becomes summation-order sensitive.  Inputs are regenerated
    # from the seed (tests/common.make_oracle_pair); their digests are stored so a drifting generator is noticed.
    d = make_oracle_pair(5, PAIR_W, PAIR_H)
    out["pair/sha_kf_img"] = sha(d["kf_img"])
    out["pair/sha_fr_img"] = sha(d["fr_img"])
    out["pair/sha_idepth"] = sha(d["idepth"])
    out["pair/K"] = np.array(d["pr"]["K"], np.float64)
    for l in range(5):
        out[f"pair/sha_image_L{l}"] = sha(d["okf"].get(O.IMAGE, l))
        out[f"pair/sha_gradients_L{l}"] = sha(d["okf"].get(O.GRADIENTS, l))
    out["pair/sha_maxgrad_L0"] = sha(d["okf"].get(O.MAXGRAD, 0))
    out["pair/num_mappable"] = np.array(d["okf"].num_mappable())
    out["pair/numData"] = np.array([d["oref"].num(l) for l in (1, 2, 3, 4)])
    init = np.array([0, 0, 0, 1, 0, 0, 0.0])
    res, trace = O.se3_track(d["oref"], d["ofr"], init, 2)
    out["pair/se3_frameToRef"] = np.array(res.frameToRef)
    out["pair/se3_scalars"] = np.array([res.lastResidual, res.pointUsage, res.lastGoodCount, res.lastBadCount, res.affine_a, res.affine_b])
    out["pair/se3_trace"] = np.array([[t[0], t[1], t[2], t[4]] for t in trace], np.float64)  # level, accepted, error, bufSize
    out["pair/sha_mask"] = sha(d["ofr"].get(O.MASK, 1))
    pres, ptrace = O.se3_track_permaref(d["oref"], d["ofr"], init, 2)
    out["pair/permaref_refToFrame"] = np.array(pres.frameToRef)
    out["pair/permaref_overlap"] = np.array(O.check_permaref_overlap(d["oref"], np.array(pres.frameToRef)))
    # ---- Sim3 tracker
    s = make_sim3_pair(O, 8, PAIR_W, PAIR_H, c=0.95)
    sres, _ = O.sim3_track(s["oref"], s["ofr"], s["gt8"] * np.array([1, 1, 1, 1, 1, 1, 1, 1.03]), 4, 1, 2)
    out["sim3/seed_c"] = np.array([8, 0.95])
    out["sim3/frameToRef"] = np.array(sres.frameToRef)
    out["sim3/hessian_diag"] = np.array(sres.hessian).reshape(7, 7).diagonal().copy()
    # ---- DepthMap: updateKeyframe + createKeyFrame, hypothesis maps exact
    O.set_exact_sums(1)
    sc = make_oracle_depth_scene(4, w, h, n_refs=3)
    m0 = hyp_from_idepth(sc["idepth"], sc["var"])
    dm = O.DepthMap(w, h, sc["K"])
    dm.init_map(sc["okf"], m0)
    dm.update_keyframe([r["of"] for r in sc["refs"]])
    m1 = dm.read()
    out["depth/sha_map_after_update"] = sha(canonical_map(m1))
    out["depth/valid_after_update"] = np.array(int(m1["isValid"].sum()))
    out["depth/idepth_sum_after_update"] = np.array(float(m1["idepth"][m1["isValid"] > 0].astype(np.float64).sum()))
    dm.create_keyframe(sc["refs"][-1]["of"])
    m2 = dm.read()
    out["depth/sha_map_after_create"] = sha(canonical_map(m2))
    out["depth/valid_after_create"] = np.array(int(m2["isValid"].sum()))
    out["depth/rescale"] = np.array(dm.last_rescale())
    O.set_exact_sums(0)
    return out


if __name__ == "__main__":
    o = build()
    p = os.path.join(ROOT, "tests", "golden", "core_regression.npz")
    np.savez_compressed(p, **o)
    print("wrote", p, os.path.getsize(p), "bytes;", len(o), "entries")
