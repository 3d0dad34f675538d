"""Synthetic textured-room sequences (SURVEY.md 8d, config 1/2 shapes).

A camera inside a 6 x 3 x 4 m box room whose six walls carry band-limited value noise
(4 octaves).  Pinhole projection, no distortion (undistortion happens upstream of the hot
path: /root/reference/lib/App/InputThread.cpp:62).  Output is what the reference feeds
SlamSystem::nextImage: an 8-bit grey image (InputThread.cpp:59,65,71), plus ground-truth
z-depth and camera-to-world pose.  torch is used only as an array library here (CPU in the
tests, CUDA in bench.py); nothing in this file is on the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

ROOM_MIN = (-3.0, -1.5, -2.0)
ROOM_MAX = (3.0, 1.5, 2.0)
OCTAVE_CELL = (0.40, 0.16, 0.07, 0.03)   # metres per lattice cell
OCTAVE_AMP = (1.0, 0.8, 0.6, 0.45)
TABLE = 128


@dataclass
class Room:
    tables: torch.Tensor  # [6, 4, TABLE, TABLE] in [-1, 1]
    contrast: float


def make_room(seed: int, device="cpu", contrast: float = 60.0) -> Room:
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    t = torch.rand((6, len(OCTAVE_CELL), TABLE, TABLE), generator=g, dtype=torch.float32) * 2.0 - 1.0
    return Room(t.to(device), contrast)


def quat_to_R(q):
    """(x, y, z, w) unit quaternion -> 3x3 rotation (numpy, float64)."""
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def R_to_quat(R):
    """3x3 rotation -> (x, y, z, w), w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = q
    q = np.array([x, y, z, w])
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def rodrigues(rv):
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * K @ K


def pose7(R, t):
    """Sophus SE3 data layout: quaternion (x,y,z,w) then translation."""
    return np.concatenate([R_to_quat(R), np.asarray(t, dtype=np.float64)])


def pose7_to_Rt(p):
    return quat_to_R(p[:4]), np.asarray(p[4:7], dtype=np.float64)


def render(room: Room, w: int, h: int, K, R_wc, t_wc, noise_seed: int | None = 0, sigma: float = 1.0):
    """Ray-cast one frame.  Returns (u8 image [h,w], z-depth f32 [h,w]) as torch tensors on room's device."""
    dev = room.tables.device
    fx, fy, cx, cy = K
    ys, xs = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32),
                            torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
    dc = torch.stack([(xs - cx) / fx, (ys - cy) / fy, torch.ones_like(xs)], dim=-1)  # [h,w,3]
    R = torch.as_tensor(np.asarray(R_wc), dtype=torch.float32, device=dev)
    o = torch.as_tensor(np.asarray(t_wc), dtype=torch.float32, device=dev)
    d = dc @ R.T
    lo = torch.tensor(ROOM_MIN, device=dev)
    hi = torch.tensor(ROOM_MAX, device=dev)
    bound = torch.where(d > 0, hi, lo)
    tt = (bound - o) / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
    tt = torch.where(d.abs() < 1e-9, torch.full_like(tt, 1e9), tt)
    t, axis = tt.min(dim=-1)
    p = o + t.unsqueeze(-1) * d
    sign = (torch.gather(d, -1, axis.unsqueeze(-1)).squeeze(-1) > 0).long()
    wall = axis * 2 + sign
    # texture coordinates = the two coordinates that are not `axis`
    ua = (axis + 1) % 3
    va = (axis + 2) % 3
    u = torch.gather(p, -1, ua.unsqueeze(-1)).squeeze(-1) + 7.0
    v = torch.gather(p, -1, va.unsqueeze(-1)).squeeze(-1) + 7.0
    val = torch.zeros_like(u)
    for o_i, (cell, amp) in enumerate(zip(OCTAVE_CELL, OCTAVE_AMP)):
        uu = u / cell
        vv = v / cell
        iu = torch.floor(uu)
        iv = torch.floor(vv)
        fu = uu - iu
        fv = vv - iv
        fu = fu * fu * (3 - 2 * fu)
        fv = fv * fv * (3 - 2 * fv)
        iu = iu.long() % TABLE
        iv = iv.long() % TABLE
        iu1 = (iu + 1) % TABLE
        iv1 = (iv + 1) % TABLE
        tab = room.tables[:, o_i]
        a = tab[wall, iv, iu]
        b = tab[wall, iv, iu1]
        c = tab[wall, iv1, iu]
        e = tab[wall, iv1, iu1]
        val = val + amp * ((a * (1 - fu) + b * fu) * (1 - fv) + (c * (1 - fu) + e * fu) * fv)
    img = 128.0 + room.contrast * val / sum(OCTAVE_AMP) * 2.0
    if noise_seed is not None and sigma > 0:
        g = torch.Generator(device="cpu").manual_seed(int(noise_seed) + 7919)
        img = img + (torch.randn((h, w), generator=g, dtype=torch.float32) * sigma).to(dev)
    img8 = img.round().clamp(0, 255).to(torch.uint8)
    return img8, t.to(torch.float32)


def default_K(w: int, h: int):
    """fx=fy=525, cx=319.5, cy=239.5 at 640x480, scaled with width."""
    s = w / 640.0
    return (525.0 * s, 525.0 * s, (w - 1) / 2.0, (h - 1) / 2.0)


def d2_K():
    """1280x960 intrinsics scaled from /root/reference/d2_camera.xml:4-8."""
    return (953.4, 953.4, 623.4, 495.1)


def random_camera(rng: np.random.Generator):
    """A camera-to-world pose well inside the room looking roughly at a corner region."""
    t = np.array([rng.uniform(-1.0, 1.0), rng.uniform(-0.4, 0.4), rng.uniform(-0.8, 0.2)])
    yaw = rng.uniform(-0.6, 0.6)
    pitch = rng.uniform(-0.2, 0.2)
    roll = rng.uniform(-0.1, 0.1)
    R = rodrigues([0, yaw, 0]) @ rodrigues([pitch, 0, 0]) @ rodrigues([0, 0, roll])
    return R, t


def small_motion(rng: np.random.Generator, max_t=0.03, max_r=math.radians(1.0)):
    """Relative motion (R, t) with |t| <= max_t and angle <= max_r, uniformly scaled."""
    tv = rng.normal(size=3)
    tv = tv / np.linalg.norm(tv) * rng.uniform(0.2, 1.0) * max_t
    rv = rng.normal(size=3)
    rv = rv / np.linalg.norm(rv) * rng.uniform(0.2, 1.0) * max_r
    return rodrigues(rv), tv


def make_pair(seed: int, w: int, h: int, K=None, device="cpu", max_t=0.03, max_r=math.radians(1.0), sigma=1.0):
    """Config-2 style (keyframe, frame) pair.

    Returns dict with kf_img, kf_depth, fr_img, fr_depth (torch), frameToRef (pose7, numpy) = GT.
    """
    K = K or default_K(w, h)
    rng = np.random.default_rng(seed)
    room = make_room(seed, device)
    R0, t0 = random_camera(rng)
    dR, dt = small_motion(rng, max_t, max_r)
    # frame pose: T_w_f = T_w_kf * T_kf_f   with T_kf_f = (dR, dt) = frameToRef
    R1 = R0 @ dR
    t1 = t0 + R0 @ dt
    kf_img, kf_depth = render(room, w, h, K, R0, t0, noise_seed=2 * seed, sigma=sigma)
    fr_img, fr_depth = render(room, w, h, K, R1, t1, noise_seed=2 * seed + 1, sigma=sigma)
    return dict(kf_img=kf_img, kf_depth=kf_depth, fr_img=fr_img, fr_depth=fr_depth, frameToRef=pose7(dR, dt), K=K,
                R_w_kf=R0, t_w_kf=t0, R_w_f=R1, t_w_f=t1, room=room)


def trajectory(n: int, seed: int = 0, step_t=0.02, step_r=math.radians(0.5)):
    """Smooth sinusoidal trajectory (config 1): list of camera-to-world (R, t)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * math.pi, size=6)
    out = []
    for i in range(n):
        s = i / 50.0
        t = np.array([0.45 * math.sin(s * 0.9 + ph[0]), 0.15 * math.sin(s * 1.3 + ph[1]), -0.5 + 0.3 * math.sin(s * 0.7 + ph[2])])
        rv = np.array([0.06 * math.sin(s * 1.1 + ph[3]), 0.25 * math.sin(s * 0.6 + ph[4]), 0.04 * math.sin(s * 0.8 + ph[5])])
        out.append((rodrigues(rv), t))
    return out


def semidense_idepth(kf_depth: torch.Tensor, max_grad: np.ndarray, var: float = 0.01, noise: float = 0.0, seed: int = 0):
    """What a converged DepthMap hands Frame::setDepth: idepth where maxGrad >= 5, else (-1,-1)."""
    d = kf_depth.detach().cpu().numpy().astype(np.float32)
    idepth = (1.0 / d).astype(np.float32)
    if noise > 0:
        rng = np.random.default_rng(seed + 101)
        idepth = (idepth * (1.0 + rng.normal(size=idepth.shape).astype(np.float32) * noise)).astype(np.float32)
    valid = max_grad >= 5.0
    valid[:3, :] = False
    valid[-3:, :] = False
    valid[:, :3] = False
    valid[:, -3:] = False
    idv = np.where(valid, idepth, np.float32(-1)).astype(np.float32)
    vv = np.where(valid, np.float32(var), np.float32(-1)).astype(np.float32)
    return idv, vv


def make_depth_scene(seed: int, w: int, h: int, n_refs: int = 10, K=None, device="cpu", step=0.01, sigma=1.0,
                     rot_step=math.radians(0.15)):
    """Config-3 style scene: one keyframe plus n_refs reference frames on a smooth path with
    baselines step, 2*step, ... (metres) and a slowly growing rotation.

    Returns dict: kf_img/kf_depth (torch), refs = list of dict(img, depth, toKf = pose7 refToKf (GT)), K.
    """
    K = K or default_K(w, h)
    rng = np.random.default_rng(seed)
    room = make_room(seed, device)
    R0, t0 = random_camera(rng)
    dirv = rng.normal(size=3)
    dirv[2] *= 0.3  # mostly sideways motion: good stereo baselines
    dirv /= np.linalg.norm(dirv)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    kf_img, kf_depth = render(room, w, h, K, R0, t0, noise_seed=1000 * seed, sigma=sigma)
    refs = []
    for i in range(1, n_refs + 1):
        dR = rodrigues(axis * rot_step * i)
        dt = dirv * step * i
        R1 = R0 @ dR
        t1 = t0 + R0 @ dt
        img, depth = render(room, w, h, K, R1, t1, noise_seed=1000 * seed + i, sigma=sigma)
        refs.append(dict(img=img, depth=depth, toKf=pose7(dR, dt), R_w=R1, t_w=t1))
    return dict(kf_img=kf_img, kf_depth=kf_depth, refs=refs, K=K, R_w_kf=R0, t_w_kf=t0, room=room)


def make_constraint_scene(seed: int, w: int, h: int, n_cand: int = 64, K=None, device="cpu", max_t=0.30, max_r=math.radians(10.0),
                          sigma=1.0):
    """Config-4 scene: ONE new keyframe and n_cand candidate keyframes of the same room within max_t / max_r
    of it (upstream: the candidates TrackableKeyFrameSearch hands findConstraintsForNewKeyFrames).

    Returns dict: new = (img, depth), cands = list of dict(img, depth, candToNew = pose7 GT candidate -> new keyframe), K.
    """
    K = K or default_K(w, h)
    rng = np.random.default_rng(seed)
    room = make_room(seed, device)
    R0, t0 = random_camera(rng)
    new_img, new_depth = render(room, w, h, K, R0, t0, noise_seed=7000 * seed, sigma=sigma)
    cands = []
    for i in range(n_cand):
        dR, dt = small_motion(rng, max_t, max_r)  # T_new_cand: candidate -> new keyframe
        R1 = R0 @ dR
        t1 = t0 + R0 @ dt
        img, depth = render(room, w, h, K, R1, t1, noise_seed=7000 * seed + 1 + i, sigma=sigma)
        cands.append(dict(img=img, depth=depth, candToNew=pose7(dR, dt)))
    return dict(new=(new_img, new_depth), cands=cands, K=K, room=room)
