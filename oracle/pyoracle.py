"""ctypes binding of the CPU oracle (oracle/_build/liblsd_oracle*.so).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/lsd_oracle.hpp).  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

IMAGE, GRADIENTS, MAXGRAD, IDEPTH, IDEPTHVAR, MASK = range(6)


class SE3Result(C.Structure):
    _fields_ = [("frameToRef", C.c_double * 7),
                ("lastResidual", C.c_float), ("lastMeanRes", C.c_float), ("pointUsage", C.c_float),
                ("lastGoodCount", C.c_float), ("lastBadCount", C.c_float),
                ("affine_a", C.c_float), ("affine_b", C.c_float), ("initialTrackedResidual", C.c_float),
                ("diverged", C.c_int), ("trackingWasGood", C.c_int),
                ("numResidualCalls", C.c_int * 5), ("numWarpUpdateCalls", C.c_int * 5),
                ("traceLen", C.c_int)]


class TraceEntry(C.Structure):
    _fields_ = [("level", C.c_int), ("accepted", C.c_int), ("error", C.c_float), ("lam", C.c_float), ("bufSize", C.c_int)]


class Sim3Result(C.Structure):
    _fields_ = [("frameToRef", C.c_double * 8),
                ("hessian", C.c_float * 49),
                ("lastResidual", C.c_float), ("lastDepthResidual", C.c_float), ("lastPhotometricResidual", C.c_float),
                ("pointUsage", C.c_float), ("affine_a", C.c_float), ("affine_b", C.c_float),
                ("diverged", C.c_int), ("traceLen", C.c_int)]


class Hypothesis(C.Structure):
    _fields_ = [("isValid", C.c_bool), ("blacklisted", C.c_int), ("nextStereoFrameMinID", C.c_float),
                ("validity_counter", C.c_int), ("idepth", C.c_float), ("idepth_var", C.c_float),
                ("idepth_smoothed", C.c_float), ("idepth_var_smoothed", C.c_float)]


HYP_DTYPE = np.dtype([("isValid", np.uint8), ("_pad", np.uint8, (3,)), ("blacklisted", np.int32),
                      ("nextStereoFrameMinID", np.float32), ("validity_counter", np.int32), ("idepth", np.float32),
                      ("idepth_var", np.float32), ("idepth_smoothed", np.float32), ("idepth_var_smoothed", np.float32)])
assert HYP_DTYPE.itemsize == 32

_libs = {}


def build(force: bool = False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    out = os.path.join(HERE, "_build", "liblsd_oracle.so")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return out


def lib(fast: bool = False):
    key = "fast" if fast else "parity"
    if key in _libs:
        return _libs[key]
    path = os.path.join(HERE, "_build", "liblsd_oracle_fast.so" if fast else "liblsd_oracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    vp, ip, fp, dp = C.c_void_p, C.c_int, C.c_float, C.c_double
    L.lsdo_frame_create.restype = vp
    L.lsdo_frame_create.argtypes = [ip, ip, ip, fp, fp, fp, fp, vp]
    L.lsdo_frame_destroy.argtypes = [vp]
    L.lsdo_frame_build_pyramids.argtypes = [vp]
    L.lsdo_frame_get.argtypes = [vp, ip, ip, vp]
    L.lsdo_frame_num_mappable.argtypes = [vp]
    L.lsdo_frame_mean_idepth.argtypes = [vp]
    L.lsdo_frame_mean_idepth.restype = fp
    L.lsdo_frame_num_points.argtypes = [vp]
    L.lsdo_frame_set_idepth.argtypes = [vp, vp, vp]
    L.lsdo_frame_set_depth_gt.argtypes = [vp, vp, fp]
    L.lsdo_frame_set_track_meta.argtypes = [vp, fp, ip, vp]
    L.lsdo_frame_set_mask.argtypes = [vp, vp]
    L.lsdo_frame_set_counters.argtypes = [vp, ip, ip]
    L.lsdo_ref_create.restype = vp
    L.lsdo_ref_create.argtypes = [vp]
    L.lsdo_ref_destroy.argtypes = [vp]
    L.lsdo_ref_num.argtypes = [vp, ip]
    L.lsdo_ref_get.argtypes = [vp, ip, vp, vp, vp, vp]
    L.lsdo_se3_track.argtypes = [vp, vp, vp, ip, vp, vp, ip]
    L.lsdo_se3_track_permaref.argtypes = [vp, vp, vp, ip, vp, vp, ip]
    L.lsdo_check_permaref_overlap.restype = fp
    L.lsdo_check_permaref_overlap.argtypes = [vp, vp]
    L.lsdo_se3_eval.argtypes = [vp, vp, vp, ip, fp, fp, ip, vp, vp, vp]
    L.lsdo_se3_track_batch.restype = dp
    L.lsdo_se3_track_batch.argtypes = [ip, vp, vp, vp, ip, ip, vp]
    L.lsdo_hardware_threads.restype = ip
    L.lsdo_make_pairs.argtypes = [ip, vp, vp, vp, vp, ip, ip, vp, ip, vp, vp, vp]
    L.lsdo_pairs_trim.restype = None
    L.lsdo_pairs_trim.argtypes = [ip, vp, vp]
    for name, res, args in [
        ("lsdo_sim3_track", ip, [vp, vp, vp, ip, ip, ip, vp, vp, ip]),
        ("lsdo_sim3_track_batch", dp, [ip, vp, vp, vp, ip, ip, ip, ip, vp]),
        ("lsdo_depthmap_create", vp, [ip, ip, fp, fp, fp, fp, ip]),
        ("lsdo_depthmap_destroy", None, [vp]),
        ("lsdo_depthmap_set_thresholds", None, [vp, ip, ip, ip, ip]),
        ("lsdo_depthmap_init_gt", None, [vp, vp]),
        ("lsdo_depthmap_init_random", None, [vp, vp, C.c_uint]),
        ("lsdo_depthmap_init_map", None, [vp, vp, vp]),
        ("lsdo_depthmap_set_reactivated", None, [vp, ip]),
        ("lsdo_depthmap_read", None, [vp, vp]),
        ("lsdo_depthmap_write", None, [vp, vp]),
        ("lsdo_depthmap_last_rescale", fp, [vp]),
        ("lsdo_depthmap_update_keyframe", dp, [vp, ip, vp, vp]),
        ("lsdo_depthmap_create_keyframe", dp, [vp, vp]),
        ("lsdo_depthmap_finalize", None, [vp]),
        ("lsdo_depthmap_stage", dp, [vp, ip, ip, ip, vp]),
        ("lsdo_depthmap_prepare", None, [vp, ip, vp]),
        ("lsdo_depthmap_debug_rgb", None, [vp, vp]),
        ("lsdo_line_stereo", fp, [vp, vp, fp, fp, fp, fp, fp, fp, fp, vp]),
        ("lsdo_make_epl", ip, [vp, vp, ip, ip, vp]),
        ("lsdo_frame_get_pose", None, [vp, vp]),
        ("lsdo_frame_get_counters", None, [vp, vp]),
        ("lsdo_frame_set_flags", None, [vp, ip]),
        ("lsdo_frame_clear_mask", None, [vp]),
        ("lsdo_frame_get_flags", ip, [vp]),
        ("lsdo_set_exact_sums", None, [ip]),
        ("lsdo_init_undistort_rectify_map", None, [vp, vp, vp, ip, ip, vp, vp]),
        ("lsdo_remap_u8", None, [vp, ip, ip, C.c_size_t, vp, vp, ip, ip, vp]),
        ("lsdo_publish_keyframe_pack", None, [vp, vp, vp, ip, ip, vp]),
        ("lsdo_compute_vbo", ip, [vp, ip, ip, fp, fp, fp, fp, fp, fp, fp, ip, ip, vp]),
    ]:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _libs[key] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Frame:
    def __init__(self, fid, img_u8: np.ndarray, K, fast=False):
        self.L = lib(fast)
        img_u8 = np.ascontiguousarray(img_u8, dtype=np.uint8)
        self.h, self.w = img_u8.shape
        self.K = K
        self.p = self.L.lsdo_frame_create(fid, self.w, self.h, K[0], K[1], K[2], K[3], _ptr(img_u8))
        self.id = fid

    def __del__(self):
        try:
            self.L.lsdo_frame_destroy(self.p)
        except Exception:
            pass

    def build_pyramids(self):
        self.L.lsdo_frame_build_pyramids(self.p)

    def get(self, field, level):
        w, h = self.w >> level, self.h >> level
        if field == GRADIENTS:
            out = np.empty((h, w, 4), np.float32)
        elif field == MASK:
            out = np.empty((self.h >> 1, self.w >> 1), np.uint8)
        else:
            out = np.empty((h, w), np.float32)
        rc = self.L.lsdo_frame_get(self.p, field, level, _ptr(out))
        if rc != 0:
            raise RuntimeError(f"lsdo_frame_get rc={rc}")
        return out

    def num_mappable(self):
        return self.L.lsdo_frame_num_mappable(self.p)

    def set_idepth(self, idepth, var):
        idepth = np.ascontiguousarray(idepth, np.float32)
        var = np.ascontiguousarray(var, np.float32)
        self.L.lsdo_frame_set_idepth(self.p, _ptr(idepth), _ptr(var))

    def set_depth_gt(self, depth, cov=1.0):
        depth = np.ascontiguousarray(depth, np.float32)
        self.L.lsdo_frame_set_depth_gt(self.p, _ptr(depth), cov)

    def set_track_meta(self, initialTrackedResidual, parent_id, toParent8):
        a = np.ascontiguousarray(toParent8, np.float64)
        self.L.lsdo_frame_set_track_meta(self.p, float(initialTrackedResidual), int(parent_id), _ptr(a))

    def set_mask(self, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        self.L.lsdo_frame_set_mask(self.p, _ptr(m))

    def set_counters(self, tracked, mapped):
        self.L.lsdo_frame_set_counters(self.p, tracked, mapped)

    def mean_idepth(self):
        return self.L.lsdo_frame_mean_idepth(self.p)

    def num_points(self):
        return self.L.lsdo_frame_num_points(self.p)


class Ref:
    def __init__(self, kf: Frame):
        self.L = kf.L
        self.kf = kf
        self.p = self.L.lsdo_ref_create(kf.p)

    def __del__(self):
        try:
            self.L.lsdo_ref_destroy(self.p)
        except Exception:
            pass

    def num(self, level):
        return self.L.lsdo_ref_num(self.p, level)

    def get(self, level):
        n = self.num(level)
        pos = np.empty((n, 3), np.float32)
        grad = np.empty((n, 2), np.float32)
        cv = np.empty((n, 2), np.float32)
        idx = np.empty((n,), np.int32)
        self.L.lsdo_ref_get(self.p, level, _ptr(pos), _ptr(grad), _ptr(cv), _ptr(idx))
        return pos, grad, cv, idx


def se3_track(ref: Ref, frame: Frame, init7, mode=0, trace_cap=2048):
    res = SE3Result()
    tr = (TraceEntry * trace_cap)()
    init = np.ascontiguousarray(init7, np.float64)
    ref.L.lsdo_se3_track(ref.p, frame.p, _ptr(init), mode, C.byref(res), tr, trace_cap)
    trace = [(tr[i].level, tr[i].accepted, tr[i].error, tr[i].lam, tr[i].bufSize) for i in range(min(res.traceLen, trace_cap))]
    return res, trace


def se3_track_permaref(ref: Ref, frame: Frame, init_ref_to_frame7, mode=0, trace_cap=2048):
    """SE3Tracker::trackFrameOnPermaref (test-track settings); the result's frameToRef holds referenceToFrame."""
    res = SE3Result()
    tr = (TraceEntry * trace_cap)()
    init = np.ascontiguousarray(init_ref_to_frame7, np.float64)
    ref.L.lsdo_se3_track_permaref(ref.p, frame.p, _ptr(init), mode, C.byref(res), tr, trace_cap)
    trace = [(tr[i].level, tr[i].accepted, tr[i].error, tr[i].lam, tr[i].bufSize) for i in range(min(res.traceLen, trace_cap))]
    return res, trace


def check_permaref_overlap(ref: Ref, ref_to_frame7):
    p = np.ascontiguousarray(ref_to_frame7, np.float64)
    return float(ref.L.lsdo_check_permaref_overlap(ref.p, _ptr(p)))


def se3_eval(ref: Ref, frame: Frame, refToFrame7, level, a=1.0, b=0.0, mode=0):
    A = np.zeros((6, 6), np.float32)
    bb = np.zeros(6, np.float32)
    sc = np.zeros(12, np.float32)
    p = np.ascontiguousarray(refToFrame7, np.float64)
    ref.L.lsdo_se3_eval(ref.p, frame.p, _ptr(p), level, a, b, mode, _ptr(A), _ptr(bb), _ptr(sc))
    return A, bb, sc


def se3_track_batch(refs, frames, inits, mode=0, threads=1):
    n = len(refs)
    L = refs[0].L
    rp = (C.c_void_p * n)(*[r.p for r in refs])
    fp = (C.c_void_p * n)(*[f.p for f in frames])
    init = np.ascontiguousarray(inits, np.float64).reshape(n, 7)
    outs = (SE3Result * n)()
    secs = L.lsdo_se3_track_batch(n, rp, fp, _ptr(init), mode, threads, outs)
    return secs, outs


class RawBatch:
    """n prepared (ref, frame) pairs living in the oracle library (bench.py cpu_baseline / --impl reference)."""

    def __init__(self, kf_imgs, fr_imgs, idepths, vars_, K, threads, fast=True, trim=False):
        self.L = lib(fast)
        n = len(kf_imgs)
        self.n = n
        self._trim = trim
        h, w = kf_imgs[0].shape
        keep = [np.ascontiguousarray(a) for a in kf_imgs], [np.ascontiguousarray(a) for a in fr_imgs], \
               [np.ascontiguousarray(a, np.float32) for a in idepths], [np.ascontiguousarray(a, np.float32) for a in vars_]
        arr = [(C.c_void_p * n)(*[a.ctypes.data for a in lst]) for lst in keep]
        self.kf = (C.c_void_p * n)()
        self.fr = (C.c_void_p * n)()
        self.ref = (C.c_void_p * n)()
        Kc = (C.c_float * 4)(*K)
        self.L.lsdo_make_pairs(n, arr[0], arr[1], arr[2], arr[3], w, h, Kc, threads, self.kf, self.fr, self.ref)
        if trim:  # keep only what SE3Tracker::trackFrame reads (full-size batches on the CPU arm)
            self.L.lsdo_pairs_trim(n, self.kf, self.fr)

    def track(self, inits, mode=0, threads=1):
        init = np.ascontiguousarray(inits, np.float64).reshape(self.n, 7)
        outs = (SE3Result * self.n)()
        secs = self.L.lsdo_se3_track_batch(self.n, self.ref, self.fr, _ptr(init), mode, threads, outs)
        return secs, outs

    def free(self):
        for i in range(self.n):
            self.L.lsdo_ref_destroy(self.ref[i])
            self.L.lsdo_frame_destroy(self.kf[i])
            self.L.lsdo_frame_destroy(self.fr[i])
        self.n = 0


def sim3_track(ref: Ref, frame: Frame, init8, start_level=4, final_level=1, mode=0, trace_cap=2048):
    """Sim3Tracker::trackFrameSim3; init8 = frameToReference (qx,qy,qz,qw,tx,ty,tz,scale)."""
    res = Sim3Result()
    tr = (TraceEntry * trace_cap)()
    init = np.ascontiguousarray(init8, np.float64)
    ref.L.lsdo_sim3_track(ref.p, frame.p, _ptr(init), start_level, final_level, mode, C.byref(res), tr, trace_cap)
    trace = [(tr[i].level, tr[i].accepted, tr[i].error, tr[i].lam, tr[i].bufSize) for i in range(min(res.traceLen, trace_cap))]
    return res, trace


def sim3_track_batch(refs, frames, inits, start_level=4, final_level=1, mode=0, threads=1):
    n = len(refs)
    L = refs[0].L
    rp = (C.c_void_p * n)(*[r.p for r in refs])
    fp = (C.c_void_p * n)(*[f.p for f in frames])
    init = np.ascontiguousarray(inits, np.float64).reshape(n, 8)
    outs = (Sim3Result * n)()
    secs = L.lsdo_sim3_track_batch(n, rp, fp, _ptr(init), start_level, final_level, mode, threads, outs)
    return secs, outs


STAGE_OBSERVE, STAGE_FILL_HOLES, STAGE_REGULARIZE, STAGE_PROPAGATE, STAGE_SET_DEPTH = range(5)


class DepthMap:
    """[UP] lsd_slam::DepthMap restated on the CPU (oracle/depthmap.cpp)."""

    def __init__(self, w, h, K, threads=1, fast=False):
        self.L = lib(fast)
        self.w, self.h = w, h
        self.p = self.L.lsdo_depthmap_create(w, h, K[0], K[1], K[2], K[3], threads)
        self._keep = []

    def __del__(self):
        try:
            self.L.lsdo_depthmap_destroy(self.p)
        except Exception:
            pass

    def set_thresholds(self, create=30, keep=24, unblacklist=100, min_blacklist=-1):
        self.L.lsdo_depthmap_set_thresholds(self.p, create, keep, unblacklist, min_blacklist)

    def init_gt(self, frame: Frame):
        self._keep.append(frame)
        self.L.lsdo_depthmap_init_gt(self.p, frame.p)

    def init_random(self, frame: Frame, seed=1):
        self._keep.append(frame)
        self.L.lsdo_depthmap_init_random(self.p, frame.p, seed)

    def init_map(self, frame: Frame, hyp):
        self._keep.append(frame)
        hyp = np.ascontiguousarray(hyp, HYP_DTYPE)
        self.L.lsdo_depthmap_init_map(self.p, frame.p, _ptr(hyp))

    def set_reactivated(self, v):
        self.L.lsdo_depthmap_set_reactivated(self.p, int(v))

    def read(self):
        out = np.zeros((self.h, self.w), HYP_DTYPE)
        self.L.lsdo_depthmap_read(self.p, _ptr(out))
        return out

    def write(self, hyp):
        hyp = np.ascontiguousarray(hyp, HYP_DTYPE)
        self.L.lsdo_depthmap_write(self.p, _ptr(hyp))

    def update_keyframe(self, frames):
        self._keep.extend(frames)
        fp = (C.c_void_p * len(frames))(*[f.p for f in frames])
        return self.L.lsdo_depthmap_update_keyframe(self.p, len(frames), fp, None)

    def prepare(self, frames):
        self._keep.extend(frames)
        fp = (C.c_void_p * len(frames))(*[f.p for f in frames])
        self.L.lsdo_depthmap_prepare(self.p, len(frames), fp)

    def create_keyframe(self, new_kf: Frame):
        self._keep.append(new_kf)
        return self.L.lsdo_depthmap_create_keyframe(self.p, new_kf.p)

    def finalize(self):
        self.L.lsdo_depthmap_finalize(self.p)

    def stage(self, stage, arg1=0, arg2=0, frame=None):
        if frame is not None:
            self._keep.append(frame)
        return self.L.lsdo_depthmap_stage(self.p, stage, arg1, arg2, frame.p if frame is not None else None)

    def last_rescale(self):
        return self.L.lsdo_depthmap_last_rescale(self.p)

    def debug_rgb(self):
        out = np.zeros((self.h, self.w, 3), np.uint8)
        self.L.lsdo_depthmap_debug_rgb(self.p, _ptr(out))
        return out

    def line_stereo(self, ref: Frame, u, v, epxn, epyn, min_id, prior_id, max_id):
        out = np.zeros(3, np.float32)
        e = self.L.lsdo_line_stereo(self.p, ref.p, u, v, epxn, epyn, min_id, prior_id, max_id, _ptr(out))
        return e, out

    def make_epl(self, ref: Frame, x, y):
        out = np.zeros(2, np.float32)
        ok = self.L.lsdo_make_epl(self.p, ref.p, x, y, _ptr(out))
        return bool(ok), out


def frame_pose(frame: Frame):
    out = np.zeros(8)
    frame.L.lsdo_frame_get_pose(frame.p, _ptr(out))
    return out


def frame_counters(frame: Frame):
    out = np.zeros(3, np.int32)
    frame.L.lsdo_frame_get_counters(frame.p, _ptr(out))
    return out


def set_exact_sums(v, fast=False):
    """fp64 accumulation of the depth map's two whole-map sums (see lsd_oracle.hpp g_exactSums)."""
    lib(fast).lsdo_set_exact_sums(int(v))


# ---- keyframe publish / VBO extraction (oracle/keyframe.cpp; the only PINNED part of the oracle) ----------------
POINT_DTYPE = np.dtype([("idepth", np.float32), ("idepth_var", np.float32), ("color", np.uint8, (4,))])  # InputPointDense
VERTEX_DTYPE = np.dtype([("point", np.float32, (3,)), ("color", np.uint8, (4,))])  # Keyframe::MyVertex
assert POINT_DTYPE.itemsize == 12 and VERTEX_DTYPE.itemsize == 16
VBO_SCALED_TH, VBO_ABS_TH, VBO_MIN_NEAR_SUPPORT = 1e-3, 1e-1, 9  # Keyframe.h:79-83


def publish_keyframe_pack(idepth, idepth_var, image, has_idepth=True, fast=False):
    """PangolinOutputIOWrapper::publishKeyframe's pack loop (PangolinOutputIOWrapper.cpp:69-89), restated."""
    idepth = np.ascontiguousarray(idepth, np.float32)
    idepth_var = np.ascontiguousarray(idepth_var, np.float32)
    image = np.ascontiguousarray(image, np.float32)
    out = np.zeros(idepth.shape, POINT_DTYPE)
    lib(fast).lsdo_publish_keyframe_pack(_ptr(idepth), _ptr(idepth_var), _ptr(image), idepth.size, int(has_idepth), _ptr(out))
    return out


def compute_vbo(points, K, scale=1.0, scaled_th=VBO_SCALED_TH, abs_th=VBO_ABS_TH, min_near_support=VBO_MIN_NEAR_SUPPORT,
                contract_fma=False, fast=False):
    """Keyframe::computeVbo (Keyframe.h:66-158), restated.  points: (h, w) POINT_DTYPE.  Returns the vertex array."""
    points = np.ascontiguousarray(points, POINT_DTYPE)
    h, w = points.shape
    out = np.zeros(h * w, VERTEX_DTYPE)
    n = lib(fast).lsdo_compute_vbo(_ptr(points), w, h, K[0], K[1], K[2], K[3], scale, scaled_th, abs_th, min_near_support,
                                   int(contract_fma), _ptr(out))
    return out[:n].copy()


_ref_kf = None


def ref_keyframe_lib():
    """oracle/_ref/libref_keyframe.so: the reference's own Keyframe.h compiled by oracle/Makefile (None when neither
    /root/reference nor a prebuilt library is present)."""
    global _ref_kf
    if _ref_kf is None:
        path = os.path.join(HERE, "_ref", "libref_keyframe.so")
        if not os.path.exists(path):
            build(force=True)
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        vp, ip, fp = C.c_void_p, C.c_int, C.c_float
        L.ref_keyframe_compute_vbo.restype = ip
        L.ref_keyframe_compute_vbo.argtypes = [vp, ip, ip, fp, fp, fp, fp, fp, vp]
        L.ref_keyframe_update_and_compute_vbo.restype = ip
        L.ref_keyframe_update_and_compute_vbo.argtypes = [vp, vp, ip, ip, fp, fp, fp, fp, fp, vp]
        _ref_kf = L
    return _ref_kf


def ref_compute_vbo(points, K, scale=1.0, republish=None):
    """The REFERENCE's Keyframe::computeVbo itself (republish: a second point set pushed through
    Keyframe::updatePoints first, the path of GUI::addKeyframe for a known id, lib/GUI.cpp:126-131)."""
    L = ref_keyframe_lib()
    assert L is not None, "oracle/_ref/libref_keyframe.so unavailable"
    points = np.ascontiguousarray(points, POINT_DTYPE)
    h, w = points.shape
    out = np.zeros(h * w, VERTEX_DTYPE)
    if republish is None:
        n = L.ref_keyframe_compute_vbo(_ptr(points), w, h, K[0], K[1], K[2], K[3], scale, _ptr(out))
    else:
        second = np.ascontiguousarray(republish, POINT_DTYPE)
        n = L.ref_keyframe_update_and_compute_vbo(_ptr(points), _ptr(second), w, h, K[0], K[1], K[2], K[3], scale, _ptr(out))
    assert n >= 0
    return out[:n].copy()


# ---- undistortion (oracle/undistort.cpp; pinned on cv2's own output, tests/golden/undistort.npz) ------------------
def init_undistort_rectify_map(K, dist, K_out, w, h, fast=False):
    """cv::initUndistortRectifyMap(K, dist, I, K_out, (w, h), CV_16SC2), restated.  K = (fx, fy, cx, cy); dist = (k1, k2, p1, p2[, k3])."""
    Kd = np.ascontiguousarray(K, np.float64)
    dd = np.zeros(5)
    dd[:len(dist)] = dist
    Ko = np.ascontiguousarray(K_out, np.float64)
    m1 = np.zeros((h, w, 2), np.int16)
    m2 = np.zeros((h, w), np.uint16)
    lib(fast).lsdo_init_undistort_rectify_map(_ptr(Kd), _ptr(dd), _ptr(Ko), w, h, _ptr(m1), _ptr(m2))
    return m1, m2


def remap_u8(src, map1, map2, fast=False):
    """cv::remap(src, map1, map2, INTER_LINEAR) for 8-bit single-channel images, BORDER_CONSTANT 0, restated."""
    src = np.ascontiguousarray(src, np.uint8)
    map1 = np.ascontiguousarray(map1, np.int16)
    map2 = np.ascontiguousarray(map2, np.uint16)
    h, w = map2.shape
    dst = np.zeros((h, w), np.uint8)
    lib(fast).lsdo_remap_u8(_ptr(src), src.shape[1], src.shape[0], src.shape[1], _ptr(map1), _ptr(map2), w, h, _ptr(dst))
    return dst
