#!/usr/bin/env python
"""Generates tests/golden/undistort.npz from OpenCV itself (cv2, the Python build of the library whose
cv::initUndistortRectifyMap / cv::remap do the arithmetic of libvideoio::Undistorter::undistort,
/root/reference/lib/App/InputThread.cpp:62).  Inputs are stored next to the outputs; full-size cases are stored as
SHA-256 digests plus the calibration that regenerates them.  Run:  python scripts/make_golden_undistort.py
"""
import hashlib
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# /root/reference/d2_camera.xml:3-13 (Photoscan frame model: cx, cy are offsets from the image centre)
D2 = dict(w=1920, h=1080, f=1430.15016509976, cx=-24.1342060175983, cy=17.5452248329602,
          dist=(-0.118322055927498, 0.293632083518507, 0.000330080086304322, 0.00182154893322038, 0.0))


def d2_K(scale=1.0):
    w, h = D2["w"] * scale, D2["h"] * scale
    return (D2["f"] * scale, D2["f"] * scale, (w - 1) / 2 + D2["cx"] * scale, (h - 1) / 2 + D2["cy"] * scale)


def texture(seed, w, h):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = 128 + 60 * np.sin(x / 5.0 + seed) * np.cos(y / 7.0) + 40 * rng.standard_normal((h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def case(K, dist, in_wh, out_wh, seed, alpha=0.0):
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]])
    Kn, _ = cv2.getOptimalNewCameraMatrix(Km, np.array(dist), in_wh, alpha, out_wh)
    Ko = np.array([Kn[0, 0], Kn[1, 1], Kn[0, 2], Kn[1, 2]])
    m1, m2 = cv2.initUndistortRectifyMap(Km, np.array(dist), None, Kn, out_wh, cv2.CV_16SC2)
    img = texture(seed, *in_wh)
    out = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
    return dict(K=np.array(K), dist=np.array(dist), Kout=Ko, in_wh=np.array(in_wh), map1=m1, map2=m2, image=img, undistorted=out)


def main():
    cases = {}
    # small, stored in full: d2 calibration scaled to 240x135 -> 128x96 (crop) and -> 160x112 with alpha = 1 (the map
    # leaves the source image: BORDER_CONSTANT taps)
    cases["d2_small_crop"] = case(d2_K(0.125), D2["dist"], (240, 135), (128, 96), 1, 0.0)
    cases["d2_small_full"] = case(d2_K(0.125), D2["dist"], (240, 135), (160, 112), 2, 1.0)
    strong = (-0.35, 0.15, 0.002, -0.001, -0.03)
    cases["strong_k3"] = case((110.0, 112.0, 79.3, 59.1), strong, (160, 120), (96, 80), 3, 0.5)
    flat = {f"{n}/{k}": v for n, c in cases.items() for k, v in c.items()}
    # full size: d2 camera 1920x1080 -> 1280x960 and -> 640x480, digests only
    for name, out_wh in (("d2_1280x960", (1280, 960)), ("d2_640x480", (640, 480))):
        c = case(d2_K(), D2["dist"], (1920, 1080), out_wh, 7, 0.0)
        for k in ("K", "dist", "Kout", "in_wh"):
            flat[f"{name}/{k}"] = c[k]
        flat[f"{name}/out_wh"] = np.array(out_wh)
        flat[f"{name}/seed"] = np.array(7)
        for k in ("map1", "map2", "image", "undistorted"):
            flat[f"{name}/sha256_{k}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(c[k]).tobytes()).digest(), np.uint8)
    out = os.path.join(ROOT, "tests", "golden", "undistort.npz")
    np.savez_compressed(out, **flat)
    print("wrote", out, os.path.getsize(out), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
