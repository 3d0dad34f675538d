#!/usr/bin/env python
"""Generates tests/golden/keyframe_vbo.npz from the REFERENCE'S OWN Keyframe::computeVbo.

Needs oracle/_ref/libref_keyframe.so, i.e. /root/reference/lib/Pangolin_IOWrapper/Keyframe.h compiled by
oracle/Makefile (reference sources are read where they lie, never copied).  The inputs are stored next to the
outputs, so the consumers (tests/test_oracle_keyframe.py, tests/test_gpu_keyframe.py) need neither the reference
nor this script's RNG.  Run:  python scripts/make_golden_keyframe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402


def scene(seed, w, h, valid=0.6, noise=0.004, var_hi=3e-3, holes=True):
    """Smooth inverse-depth surface (slanted plane + bumps) with semi-dense validity, noise and a variance spread that
    straddles both thresholds of computeVbo; invalid pixels carry Frame::setDepth's -1 / -1."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    idepth = 0.8 + 0.4 * x / w + 0.2 * np.sin(y / 9.0) + 0.1 * np.cos(x / 5.0 + seed)
    idepth = (idepth + noise * rng.standard_normal((h, w))).astype(np.float32)
    var = (rng.random((h, w)) ** 2 * var_hi).astype(np.float32)
    ok = rng.random((h, w)) < valid
    if holes:  # blocky validity so that 3x3 support is met inside blobs and violated on their rims
        blk = rng.random((h // 4 + 1, w // 4 + 1)) < valid
        ok = np.kron(blk, np.ones((4, 4), bool))[:h, :w] & (rng.random((h, w)) < 0.97)
    idepth = np.where(ok, idepth, -1).astype(np.float32)
    var = np.where(ok, var, -1).astype(np.float32)
    image = rng.integers(0, 256, (h, w)).astype(np.uint8)
    return idepth, var, image


def main():
    assert O.ref_keyframe_lib() is not None, "build oracle/_ref first (make -C oracle; needs /root/reference)"
    cases = {}
    specs = [  # name, seed, w, h, K(fx, fy, cx, cy), camToWorld scale
        ("a64x48", 1, 64, 48, (52.5, 52.5, 31.5, 23.5), 1.0),
        ("b64x48_scale7", 2, 64, 48, (52.5, 50.0, 30.0, 25.0), 7.0),   # absTH active: var*depth^4*49 > 0.1
        ("c160x112", 3, 160, 112, (131.25, 131.25, 79.5, 55.5), 0.5),
        ("d48x32_dense", 4, 48, 32, (39.0, 39.0, 23.5, 15.5), 1.3),
    ]
    for name, seed, w, h, K, scale in specs:
        idepth, var, image = scene(seed, w, h, valid=0.95 if "dense" in name else 0.6, holes="dense" not in name)
        if name == "a64x48":  # specials: NaN / inf / zero / negative-zero idepth, zero variance
            idepth[10, 10] = np.nan
            idepth[12, 20] = np.inf
            idepth[14, 30] = 0.0
            idepth[16, 40] = -0.0
            var[20, 20] = 0.0
            var[22, 22] = np.nan
        pts = O.publish_keyframe_pack(idepth, var, image.astype(np.float32))
        vtx = O.ref_compute_vbo(pts, K, scale)
        cases[name] = dict(idepth=idepth, var=var, image=image, K=np.array(K, np.float32), scale=np.float32(scale),
                           points=pts.view(np.uint8).reshape(h, w, 12), vertices=vtx.view(np.uint8).reshape(-1, 16))
        print(name, "->", len(vtx), "vertices of", w * h)
    # republish path: Keyframe::updatePoints + second computeVbo (lib/GUI.cpp:126-131)
    i1, v1, im1 = scene(11, 64, 48)
    i2, v2, im2 = scene(12, 64, 48)
    K = (52.5, 52.5, 31.5, 23.5)
    p1 = O.publish_keyframe_pack(i1, v1, im1.astype(np.float32))
    p2 = O.publish_keyframe_pack(i2, v2, im2.astype(np.float32))
    vtx = O.ref_compute_vbo(p1, K, 1.0, republish=p2)
    cases["e64x48_republished"] = dict(idepth=i2, var=v2, image=im2, K=np.array(K, np.float32), scale=np.float32(1.0),
                                       points=p2.view(np.uint8).reshape(48, 64, 12), vertices=vtx.view(np.uint8).reshape(-1, 16))
    flat = {f"{n}/{k}": v for n, c in cases.items() for k, v in c.items()}
    out = os.path.join(ROOT, "tests", "golden", "keyframe_vbo.npz")
    np.savez_compressed(out, **flat)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
