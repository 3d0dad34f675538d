"""ctypes binding of liblsd_b200.so (include/lsd_b200.h).

Thin host-side mirror of the lsd-slam core classes the reference application links
(/root/reference/tools/LSD.cpp:102, lib/App/InputThread.cpp:71): Frame, TrackingReference,
SE3Tracker.  There is NO CPU fallback: if the CUDA library is missing or no sm_100 device is
present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("LSD_B200_LIB") or os.path.join(PKG_DIR, "liblsd_b200.so")  # override: kernel-variant experiments

NL = 5
TRACE_CAP = 512
FIELD_IMAGE, FIELD_GRADIENTS, FIELD_MAXGRAD, FIELD_IDEPTH, FIELD_IDEPTHVAR, FIELD_MASK = range(6)
BUILD_TRACKING, BUILD_MAXGRAD0, BUILD_GRAD0 = 0, 1, 2


class LsdError(RuntimeError):
    pass


class TrackerSettings(C.Structure):
    _fields_ = [("lambdaSuccessFac", C.c_float), ("lambdaFailFac", C.c_float),
                ("stepSizeMin", C.c_float * NL), ("convergenceEps", C.c_float * NL),
                ("maxItsPerLvl", C.c_int * NL), ("lambdaInitial", C.c_float * NL),
                ("var_weight", C.c_float), ("huber_d", C.c_float)]


class SE3Result(C.Structure):
    _fields_ = [("frameToRef", C.c_double * 7),
                ("lastResidual", C.c_float), ("lastMeanRes", C.c_float), ("pointUsage", C.c_float),
                ("lastGoodCount", C.c_float), ("lastBadCount", C.c_float),
                ("affine_a", C.c_float), ("affine_b", C.c_float), ("initialTrackedResidual", C.c_float),
                ("diverged", C.c_int), ("trackingWasGood", C.c_int),
                ("numResidualCalls", C.c_int * NL), ("numWarpUpdateCalls", C.c_int * NL),
                ("traceLen", C.c_int)]


class Sim3Result(C.Structure):
    _fields_ = [("frameToRef", C.c_double * 8), ("lastSim3Hessian", C.c_float * 49),
                ("lastResidual", C.c_float), ("lastDepthResidual", C.c_float), ("lastPhotometricResidual", C.c_float),
                ("pointUsage", C.c_float), ("affine_a", C.c_float), ("affine_b", C.c_float),
                ("diverged", C.c_int), ("numResidualCalls", C.c_int * NL), ("numWarpUpdateCalls", C.c_int * NL),
                ("traceLen", C.c_int)]


class TraceEntry(C.Structure):
    _fields_ = [("level", C.c_int), ("accepted", C.c_int), ("error", C.c_float), ("lam", C.c_float), ("bufSize", C.c_int)]


_lib = None

# every symbol include/lsd_b200.h declares: name -> (restype, argtypes)
_vp, _ip, _fp, _dp, _sz, _u = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_uint
SYMBOLS = {
    "lsd_last_error": (C.c_char_p, []),
    "lsd_version": (_ip, []),
    "lsd_ctx_create": (_ip, [_ip, _ip, _ip, _vp, _vp, _vp]),
    "lsd_ctx_destroy": (_ip, [_vp]),
    "lsd_ctx_synchronize": (_ip, [_vp]),
    "lsd_ctx_stream": (_vp, [_vp]),
    "lsd_ctx_launch_count": (C.c_longlong, [_vp]),
    "lsd_default_tracker_settings": (_ip, [_vp]),
    "lsd_ctx_set_se3_settings": (_ip, [_vp, _vp]),
    "lsd_ctx_set_se3_work_item_records": (_ip, [_vp, _ip]),
    "lsd_ctx_set_se3_active_pairs": (_ip, [_vp, _ip]),
    "lsd_ctx_set_stencil_tma": (_ip, [_vp, _ip]),
    "lsd_ctx_set_se3_record_points": (_ip, [_vp, _ip]),
    "lsd_ctx_set_se3_live_pairs": (_ip, [_vp, _ip]),
    "lsd_ctx_set_live_tracking": (_ip, [_vp, _ip]),
    "lsd_ctx_set_se3_record_points_per_level": (_ip, [_vp, _vp]),
    "lsd_frame_create": (_ip, [_vp, _ip, _vp, _sz, _u, _vp]),
    "lsd_frame_create_batch": (_ip, [_vp, _ip, _vp, _vp, _sz, _u, _vp]),
    "lsd_frame_create_batch_device": (_ip, [_vp, _ip, _vp, _vp, _u, _vp]),
    "lsd_frame_release": (_ip, [_vp, _vp]),
    "lsd_frame_release_batch": (_ip, [_vp, _ip, _vp]),
    "lsd_frame_read": (_ip, [_vp, _vp, _ip, _ip, _vp]),
    "lsd_frame_num_mappable_pixels": (_ip, [_vp, _vp, _vp]),
    "lsd_frame_set_depth_from_gt": (_ip, [_vp, _vp, _vp, _fp]),
    "lsd_frame_set_idepth": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_frame_set_idepth_batch_device": (_ip, [_vp, _ip, _vp, _vp, _vp]),
    "lsd_frame_mean_idepth": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_frame_mean_idepth_batch": (_ip, [_vp, _ip, _vp, _vp, _vp]),
    "lsd_ref_create": (_ip, [_vp, _vp, _vp]),
    "lsd_ref_create_batch": (_ip, [_vp, _ip, _vp, _vp]),
    "lsd_ref_release": (_ip, [_vp, _vp]),
    "lsd_ref_num_data": (_ip, [_vp, _vp, _ip, _vp]),
    "lsd_ref_read": (_ip, [_vp, _vp, _ip, _vp, _vp, _vp, _vp]),
    "lsd_se3_track": (_ip, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "lsd_se3_track_batch": (_ip, [_vp, _ip, _vp, _vp, _vp, _vp, _vp]),
    "lsd_se3_track_images_batch": (_ip, [_vp, _ip, _vp, _vp, _sz, _vp, _vp]),
    "lsd_ctx_set_image_pipeline": (_ip, [_vp, _ip, _ip, _dp]),
    "lsd_default_permaref_settings": (_ip, [_vp]),
    "lsd_ctx_set_permaref_settings": (_ip, [_vp, _vp]),
    "lsd_se3_track_permaref_batch": (_ip, [_vp, _ip, _vp, _vp, _vp, _vp, _vp]),
    "lsd_se3_check_permaref_overlap_batch": (_ip, [_vp, _ip, _vp, _vp, _vp]),
    "lsd_se3_eval": (_ip, [_vp, _vp, _vp, _vp, _ip, _fp, _fp, _vp, _vp, _vp]),
    "lsd_se3_last_stats": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_ctx_set_sim3_settings": (_ip, [_vp, _vp]),
    "lsd_ctx_set_sim3_record_points": (_ip, [_vp, _ip]),
    "lsd_sim3_track": (_ip, [_vp, _vp, _vp, _vp, _ip, _ip, _vp, _vp]),
    "lsd_sim3_track_batch": (_ip, [_vp, _ip, _vp, _vp, _vp, _ip, _ip, _vp, _vp]),
    "lsd_sim3_track_stages_batch": (_ip, [_vp, _ip, _vp, _vp, _vp, _ip, _vp, _vp, _vp]),
    "lsd_frame_set_tracking_meta": (_ip, [_vp, _vp, _ip, _vp, _fp]),
    "lsd_frame_get_tracking_meta": (_ip, [_vp, _vp, _vp, _vp, _vp]),
    "lsd_frame_set_mask": (_ip, [_vp, _vp, _vp]),
    "lsd_frame_set_counters": (_ip, [_vp, _vp, _ip, _ip]),
    "lsd_frame_get_counters": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_frame_set_depth_updated_flag": (_ip, [_vp, _vp, _ip]),
    "lsd_frame_get_depth_updated_flag": (_ip, [_vp, _vp, _vp]),
    "lsd_depthmap_create": (_ip, [_vp, _vp]),
    "lsd_depthmap_destroy": (_ip, [_vp, _vp]),
    "lsd_default_depth_settings": (_ip, [_vp]),
    "lsd_depthmap_set_settings": (_ip, [_vp, _vp, _vp]),
    "lsd_depth_initialize_from_gt": (_ip, [_vp, _vp, _vp]),
    "lsd_depth_initialize_randomly": (_ip, [_vp, _vp, _vp]),
    "lsd_depth_initialize_from_map": (_ip, [_vp, _vp, _vp, _vp, _ip]),
    "lsd_depth_update_keyframe": (_ip, [_vp, _vp, _ip, _vp, _vp]),
    "lsd_depth_create_keyframe": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_depth_finalize_keyframe": (_ip, [_vp, _vp]),
    "lsd_depth_update_keyframe_batch": (_ip, [_vp, _ip, _vp, _vp]),
    "lsd_depth_create_keyframe_batch": (_ip, [_vp, _ip, _vp, _vp, _vp]),
    "lsd_depth_finalize_keyframe_batch": (_ip, [_vp, _ip, _vp]),
    "lsd_depth_read": (_ip, [_vp, _vp, _vp]),
    "lsd_depth_debug_rgb": (_ip, [_vp, _vp, _vp]),
    "lsd_depth_prepare": (_ip, [_vp, _vp, _ip, _vp, _vp]),
    "lsd_depth_stage": (_ip, [_vp, _vp, _ip, _ip, _ip, _vp]),
    "lsd_ctx_last_stage_ms": (_ip, [_vp, _vp]),
    "lsd_depth_stage_batch": (_ip, [_vp, _ip, _vp, _ip, _ip, _ip, _vp]),
    "lsd_slam_create": (_ip, [_vp, _vp]),
    "lsd_slam_ref_frame_score": (_fp, [_fp, _fp]),
    "lsd_slam_destroy": (_ip, [_vp]),
    "lsd_slam_set_keep_keyframes": (_ip, [_vp, _ip]),
    "lsd_slam_set_pipelined": (_ip, [_vp, _ip]),
    "lsd_slam_set_undistorter": (_ip, [_vp, _vp]),
    "lsd_slam_gt_depth_init": (_ip, [_vp, _ip, _vp, _sz, _vp, _vp]),
    "lsd_slam_random_init": (_ip, [_vp, _ip, _vp, _sz, _vp]),
    "lsd_slam_next_image": (_ip, [_vp, _ip, _vp, _sz, _vp]),
    "lsd_slam_next_image_batch": (_ip, [_ip, _vp, _vp, _vp, _sz, _vp]),
    "lsd_slam_current_keyframe": (_ip, [_vp, _vp, _vp]),
    "lsd_slam_counters": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_slam_stage_seconds": (_ip, [_vp, _vp]),
    "lsd_slam_pose_line": (_ip, [_vp, _vp, _sz]),
    "lsd_undistorter_create_from_maps": (_ip, [_vp, _ip, _ip, _vp, _vp, _vp]),
    "lsd_undistorter_create_opencv": (_ip, [_vp, _ip, _ip, _vp, _vp, _vp, _vp]),
    "lsd_undistorter_destroy": (_ip, [_vp, _vp]),
    "lsd_undistorter_maps": (_ip, [_vp, _vp, _vp, _vp]),
    "lsd_undistort": (_ip, [_vp, _vp, _vp, _sz, _vp]),
    "lsd_frame_create_undistorted": (_ip, [_vp, _vp, _ip, _vp, _sz, _u, _vp, _vp]),
    "lsd_frame_create_undistorted_batch": (_ip, [_vp, _vp, _ip, _vp, _vp, _sz, _u, _vp, _vp]),
    "lsd_default_vbo_params": (_ip, [_vp]),
    "lsd_frame_publish_keyframe": (_ip, [_vp, _vp, _ip, _vp]),
    "lsd_keyframe_compute_vbo": (_ip, [_vp, _vp, _ip, _fp, _vp, _vp, _vp]),
    "lsd_keyframe_compute_vbo_batch": (_ip, [_vp, _ip, _vp, _ip, _vp, _vp, _vp, _vp, _vp]),
}

# InputPointDense / Keyframe::MyVertex (/root/reference/lib/Pangolin_IOWrapper/Keyframe.h:16-21,47-51)
POINT_DTYPE = np.dtype([("idepth", np.float32), ("idepth_var", np.float32), ("color", np.uint8, (4,))])
VERTEX_DTYPE = np.dtype([("point", np.float32, (3,)), ("color", np.uint8, (4,))])


class SlamStatus(C.Structure):
    _fields_ = [("frameId", C.c_int), ("tracked", C.c_int), ("isKeyframe", C.c_int), ("numKeyframes", C.c_int),
                ("currentKeyframeId", C.c_int), ("trackingWasGood", C.c_int), ("diverged", C.c_int),
                ("pointUsage", C.c_float), ("lastResidual", C.c_float), ("keyframeScore", C.c_float),
                ("camToWorld", C.c_double * 8), ("thisToParent_raw", C.c_double * 8), ("keyframeRescale", C.c_double)]


class VboParams(C.Structure):
    _fields_ = [("scaledTH", C.c_float), ("absTH", C.c_float), ("minNearSupport", C.c_int), ("sparsifyFactor", C.c_int),
                ("contractFma", C.c_int)]

# [UP] DepthMapPixelHypothesis in upstream's 32-byte AoS layout (lsd_hypothesis)
HYP_DTYPE = np.dtype([("isValid", np.uint8), ("_pad", np.uint8, (3,)), ("blacklisted", np.int32),
                      ("nextStereoFrameMinID", np.float32), ("validity_counter", np.int32), ("idepth", np.float32),
                      ("idepth_var", np.float32), ("idepth_smoothed", np.float32), ("idepth_var_smoothed", np.float32)])
STAGE_OBSERVE, STAGE_FILL_HOLES, STAGE_REGULARIZE, STAGE_PROPAGATE, STAGE_SET_DEPTH = range(5)


class DepthSettings(C.Structure):
    _fields_ = [("valSumMinForCreate", C.c_int), ("valSumMinForKeep", C.c_int), ("valSumMinForUnblacklist", C.c_int),
                ("minBlacklist", C.c_int)]


def load():
    """dlopen liblsd_b200.so; raises loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsdError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"(liblsd_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise LsdError(f"liblsd_b200 error {rc}: {load().lsd_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One device + stream + scratch (SlamSystem's tracking / mapping threads each own one)."""

    def __init__(self, width, height, K, device=0, stream=None):
        self.L = load()
        self.w, self.h = width, height
        self.K = tuple(float(k) for k in K)
        karr = (C.c_float * 4)(*self.K)
        p = C.c_void_p()
        _chk(self.L.lsd_ctx_create(device, width, height, karr, C.c_void_p(stream or 0), C.byref(p)))
        self.p = p

    def close(self):
        if self.p:
            self.L.lsd_ctx_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _chk(self.L.lsd_ctx_synchronize(self.p))

    def launch_count(self):
        return int(self.L.lsd_ctx_launch_count(self.p))

    def default_settings(self):
        s = TrackerSettings()
        _chk(self.L.lsd_default_tracker_settings(C.byref(s)))
        return s

    def set_se3_settings(self, s):
        _chk(self.L.lsd_ctx_set_se3_settings(self.p, C.byref(s)))

    def set_stencil_tma(self, mask):
        _chk(self.L.lsd_ctx_set_stencil_tma(self.p, int(mask)))

    def set_se3_active_pairs(self, n):
        _chk(self.L.lsd_ctx_set_se3_active_pairs(self.p, int(n)))

    def set_image_pipeline(self, chunk_frames=0, streamed=-1, watchdog_seconds=0.0):
        """Scheduling of se3_track_images_batch (results never change): see lsd_ctx_set_image_pipeline."""
        _chk(self.L.lsd_ctx_set_image_pipeline(self.p, int(chunk_frames), int(streamed), float(watchdog_seconds)))

    def set_se3_record_points(self, n):
        _chk(self.L.lsd_ctx_set_se3_record_points(self.p, int(n)))

    def set_se3_record_points_per_level(self, points):
        """points[l] for level l (5 entries, entry 0 ignored; 0 = the context-wide record size)."""
        arr = (C.c_int * 5)(*[int(v) for v in points])
        _chk(self.L.lsd_ctx_set_se3_record_points_per_level(self.p, arr))

    def set_live_tracking(self, enable=True):
        _chk(self.L.lsd_ctx_set_live_tracking(self.p, int(bool(enable))))

    def set_se3_live_pairs(self, n):
        _chk(self.L.lsd_ctx_set_se3_live_pairs(self.p, int(n)))

    def set_se3_work_item_records(self, n):
        _chk(self.L.lsd_ctx_set_se3_work_item_records(self.p, int(n)))

    # ---- frames
    def create_frames(self, images, ids=None, flags=BUILD_TRACKING):
        """images: list of (h, w) uint8 arrays or one (n, h, w) array (host)."""
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        n = len(imgs)
        for im in imgs:
            assert im.shape == (self.h, self.w)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        idarr = None
        if ids is not None:
            idarr = (C.c_int * n)(*ids)
        out = (C.c_void_p * n)()
        _chk(self.L.lsd_frame_create_batch(self.p, n, idarr, ptrs, self.w, flags, out))
        return [Frame(self, out[i], ids[i] if ids is not None else i) for i in range(n)]

    def create_frames_device(self, d_ptr, n, ids=None, flags=BUILD_TRACKING):
        idarr = (C.c_int * n)(*ids) if ids is not None else None
        out = (C.c_void_p * n)()
        _chk(self.L.lsd_frame_create_batch_device(self.p, n, idarr, C.c_void_p(d_ptr), flags, out))
        return [Frame(self, out[i], ids[i] if ids is not None else i) for i in range(n)]

    def create_frame(self, image, fid=0, flags=BUILD_TRACKING):
        return self.create_frames([image], [fid], flags)[0]

    def create_refs(self, keyframes):
        n = len(keyframes)
        kp = (C.c_void_p * n)(*[k.p for k in keyframes])
        out = (C.c_void_p * n)()
        _chk(self.L.lsd_ref_create_batch(self.p, n, kp, out))
        return [Ref(self, out[i], keyframes[i]) for i in range(n)]

    def set_idepth_batch_device(self, frames, d_idepth, d_var):
        n = len(frames)
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        _chk(self.L.lsd_frame_set_idepth_batch_device(self.p, n, fp, C.c_void_p(d_idepth), C.c_void_p(d_var)))

    def create_depthmap(self):
        return DepthMap(self)

    def depth_stage_batch(self, dms, stage, arg1=0, arg2=0, frames=None):
        """One DepthMap stage on n independent maps in a single set of launches (blockIdx.z = map)."""
        n = len(dms)
        dp = (C.c_void_p * n)(*[d.p for d in dms])
        fp = (C.c_void_p * n)(*[f.p for f in frames]) if frames is not None else None
        _chk(self.L.lsd_depth_stage_batch(self.p, n, dp, stage, arg1, arg2, fp))

    def last_stage_ms(self):
        ms = C.c_float()
        _chk(self.L.lsd_ctx_last_stage_ms(self.p, C.byref(ms)))
        return ms.value

    # ---- keyframe publish (PangolinOutputIOWrapper::publishKeyframe + Keyframe::computeVbo)
    def default_vbo_params(self):
        p = VboParams()
        _chk(self.L.lsd_default_vbo_params(C.byref(p)))
        return p

    def compute_vbo_batch(self, frames, scales, level=0, params=None, read=True):
        """Keyframe::computeVbo for n keyframes in one launch.  Returns (points[n], [vertex arrays] or None)."""
        n = len(frames)
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        sc = np.ascontiguousarray(scales, np.float32).reshape(n)
        pts = np.zeros(n, np.int32)
        cap = (self.w >> level) * (self.h >> level)
        outs, dp = None, None
        if read:
            outs = [np.zeros(cap, VERTEX_DTYPE) for _ in range(n)]
            dp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        _chk(self.L.lsd_keyframe_compute_vbo_batch(self.p, n, fp, level, _ptr(sc), C.byref(params) if params is not None else None,
                                                   None, dp, _ptr(pts)))
        if read:
            outs = [o[:k] for o, k in zip(outs, pts)]
        return pts, outs

    # ---- SE3 tracking
    def se3_track_batch(self, refs, frames, inits, want_trace=False):
        n = len(refs)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        init = np.ascontiguousarray(inits, np.float64).reshape(n, 7)
        res = (SE3Result * n)()
        tr = (TraceEntry * (TRACE_CAP * n))() if want_trace else None
        _chk(self.L.lsd_se3_track_batch(self.p, n, rp, fp, _ptr(init), res, tr))
        if want_trace:
            traces = []
            for i in range(n):
                m = min(res[i].traceLen, TRACE_CAP)
                traces.append([(tr[i * TRACE_CAP + k].level, tr[i * TRACE_CAP + k].accepted, tr[i * TRACE_CAP + k].error,
                                tr[i * TRACE_CAP + k].lam, tr[i * TRACE_CAP + k].bufSize) for k in range(m)])
            return res, traces
        return res

    def se3_track_permaref_batch(self, refs, frames, inits_ref_to_frame, want_trace=False):
        """SE3Tracker::trackFrameOnPermaref for n (keyframe reference, frame) candidates; frameToRef holds referenceToFrame."""
        n = len(refs)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        init = np.ascontiguousarray(inits_ref_to_frame, np.float64).reshape(n, 7)
        res = (SE3Result * n)()
        tr = (TraceEntry * (TRACE_CAP * n))() if want_trace else None
        _chk(self.L.lsd_se3_track_permaref_batch(self.p, n, rp, fp, _ptr(init), res, tr))
        if want_trace:
            return res, [[(tr[i * TRACE_CAP + k].level, tr[i * TRACE_CAP + k].accepted, tr[i * TRACE_CAP + k].error,
                           tr[i * TRACE_CAP + k].lam, tr[i * TRACE_CAP + k].bufSize) for k in range(min(res[i].traceLen, TRACE_CAP))]
                         for i in range(n)]
        return res

    def check_permaref_overlap_batch(self, refs, ref_to_frame):
        n = len(refs)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        p = np.ascontiguousarray(ref_to_frame, np.float64).reshape(n, 7)
        out = np.zeros(n, np.float32)
        _chk(self.L.lsd_se3_check_permaref_overlap_batch(self.p, n, rp, _ptr(p), _ptr(out)))
        return out

    def prepare_batch(self, refs, frames, inits):
        """Pre-marshal a batch so repeated calls cost no Python work (bench.py)."""
        n = len(refs)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        fp = (C.c_void_p * n)(*[f.p for f in frames]) if frames is not None else None
        init = np.ascontiguousarray(inits, np.float64).reshape(n, 7)
        res = (SE3Result * n)()
        return dict(n=n, rp=rp, fp=fp, init=init, res=res)

    def se3_track_prepared(self, b):
        _chk(self.L.lsd_se3_track_batch(self.p, b["n"], b["rp"], b["fp"], _ptr(b["init"]), b["res"], None))
        return b["res"]

    def se3_track_images_prepared(self, b, ip, pitch):
        _chk(self.L.lsd_se3_track_images_batch(self.p, b["n"], b["rp"], ip, pitch, _ptr(b["init"]), b["res"]))
        return b["res"]

    def se3_track(self, ref, frame, init7, want_trace=False):
        out = self.se3_track_batch([ref], [frame], [init7], want_trace)
        if want_trace:
            return out[0][0], out[1][0]
        return out[0]

    def se3_track_images_batch(self, refs, image_ptrs, pitch, inits):
        """image_ptrs: list of host addresses (pinned memory recommended)."""
        n = len(refs)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        ip = (C.c_void_p * n)(*image_ptrs)
        init = np.ascontiguousarray(inits, np.float64).reshape(n, 7)
        res = (SE3Result * n)()
        _chk(self.L.lsd_se3_track_images_batch(self.p, n, rp, ip, pitch, _ptr(init), res))
        return res

    # ---- Sim3 tracking
    def set_sim3_record_points(self, n):
        _chk(self.L.lsd_ctx_set_sim3_record_points(self.p, int(n)))

    def set_sim3_settings(self, s):
        _chk(self.L.lsd_ctx_set_sim3_settings(self.p, C.byref(s)))

    def sim3_track_batch(self, refs, frames, inits, start_level=4, final_level=1, want_trace=False):
        n = len(refs)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        init = np.ascontiguousarray(inits, np.float64).reshape(n, 8)
        res = (Sim3Result * n)()
        tr = (TraceEntry * (TRACE_CAP * n))() if want_trace else None
        _chk(self.L.lsd_sim3_track_batch(self.p, n, rp, fp, _ptr(init), start_level, final_level, res, tr))
        if want_trace:
            traces = []
            for i in range(n):
                m = min(res[i].traceLen, TRACE_CAP)
                traces.append([(tr[i * TRACE_CAP + k].level, tr[i * TRACE_CAP + k].accepted, tr[i * TRACE_CAP + k].error,
                                tr[i * TRACE_CAP + k].lam, tr[i * TRACE_CAP + k].bufSize) for k in range(m)])
            return res, traces
        return res

    def sim3_track_stages_batch(self, refs, frames, inits, stages=((4, 3), (2, 2), (1, 1))):
        """Chained trackFrameSim3 calls per track in one launch (tryTrackSim3's level schedule).  Returns a list over stages of
        lists over tracks of Sim3Result."""
        n, k = len(refs), len(stages)
        rp = (C.c_void_p * n)(*[r.p for r in refs])
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        init = np.ascontiguousarray(inits, np.float64).reshape(n, 8)
        res = (Sim3Result * (n * k))()
        sl = (C.c_int * k)(*[int(a) for a, _ in stages])
        fl = (C.c_int * k)(*[int(b) for _, b in stages])
        _chk(self.L.lsd_sim3_track_stages_batch(self.p, n, rp, fp, _ptr(init), k, sl, fl, res))
        return [[res[s * n + i] for i in range(n)] for s in range(k)]

    def sim3_track(self, ref, frame, init8, start_level=4, final_level=1, want_trace=False):
        out = self.sim3_track_batch([ref], [frame], [init8], start_level, final_level, want_trace)
        if want_trace:
            return out[0][0], out[1][0]
        return out[0]

    def se3_eval(self, ref, frame, refToFrame7, level, a=1.0, b=0.0):
        A = np.zeros((6, 6), np.float32)
        bb = np.zeros(6, np.float32)
        sc = np.zeros(12, np.float32)
        p = np.ascontiguousarray(refToFrame7, np.float64)
        _chk(self.L.lsd_se3_eval(self.p, ref.p, frame.p, _ptr(p), level, a, b, _ptr(A), _ptr(bb), _ptr(sc)))
        return A, bb, sc

    def se3_last_stats(self):
        b = C.c_double()
        e = C.c_longlong()
        ms = C.c_float()
        _chk(self.L.lsd_se3_last_stats(self.p, C.byref(b), C.byref(e), C.byref(ms)))
        return b.value, e.value, ms.value


class Frame:
    """[UP] lsd_slam::Frame: device-resident pyramids; accessors copy to host on demand."""

    def __init__(self, ctx: Context, p, fid):
        self.ctx, self.p, self.id = ctx, p, fid

    def release(self):
        if self.p:
            self.ctx.L.lsd_frame_release(self.ctx.p, self.p)
            self.p = None

    def read(self, field, level):
        w, h = self.ctx.w >> level, self.ctx.h >> level
        if field == FIELD_GRADIENTS:
            out = np.empty((h, w, 4), np.float32)
        elif field == FIELD_MASK:
            out = np.empty((self.ctx.h >> 1, self.ctx.w >> 1), np.uint8)
        else:
            out = np.empty((h, w), np.float32)
        _chk(self.ctx.L.lsd_frame_read(self.ctx.p, self.p, field, level, _ptr(out)))
        return out

    def image(self, level=0):
        return self.read(FIELD_IMAGE, level)

    def gradients(self, level=0):
        return self.read(FIELD_GRADIENTS, level)

    def maxGradients(self, level=0):
        return self.read(FIELD_MAXGRAD, level)

    def idepth(self, level=0):
        return self.read(FIELD_IDEPTH, level)

    def idepthVar(self, level=0):
        return self.read(FIELD_IDEPTHVAR, level)

    def refPixelWasGood(self):
        return self.read(FIELD_MASK, 1)

    def publish_keyframe(self, level=0):
        """InputPointDense records of PangolinOutputIOWrapper::publishKeyframe (what Keyframe::pointData holds)."""
        out = np.zeros((self.ctx.h >> level, self.ctx.w >> level), POINT_DTYPE)
        _chk(self.ctx.L.lsd_frame_publish_keyframe(self.ctx.p, self.p, level, _ptr(out)))
        return out

    def compute_vbo(self, scale=1.0, level=0, params=None):
        """Keyframe::computeVbo: the MyVertex array the reference uploads with glBufferData."""
        out = np.zeros((self.ctx.h >> level) * (self.ctx.w >> level), VERTEX_DTYPE)
        n = C.c_int()
        _chk(self.ctx.L.lsd_keyframe_compute_vbo(self.ctx.p, self.p, level, float(scale),
                                                 C.byref(params) if params is not None else None, _ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def num_mappable_pixels(self):
        v = C.c_int()
        _chk(self.ctx.L.lsd_frame_num_mappable_pixels(self.ctx.p, self.p, C.byref(v)))
        return v.value

    def set_depth_from_gt(self, depth, cov_scale=1.0):
        d = np.ascontiguousarray(depth, np.float32)
        _chk(self.ctx.L.lsd_frame_set_depth_from_gt(self.ctx.p, self.p, _ptr(d), cov_scale))

    def set_idepth(self, idepth, var):
        a = np.ascontiguousarray(idepth, np.float32)
        b = np.ascontiguousarray(var, np.float32)
        _chk(self.ctx.L.lsd_frame_set_idepth(self.ctx.p, self.p, _ptr(a), _ptr(b)))

    def set_tracking_meta(self, parent_id, toParent8, initialTrackedResidual):
        a = np.ascontiguousarray(toParent8, np.float64)
        _chk(self.ctx.L.lsd_frame_set_tracking_meta(self.ctx.p, self.p, int(parent_id), _ptr(a), float(initialTrackedResidual)))

    def tracking_meta(self):
        pid = C.c_int()
        a = np.zeros(8)
        r = C.c_float()
        _chk(self.ctx.L.lsd_frame_get_tracking_meta(self.ctx.p, self.p, C.byref(pid), _ptr(a), C.byref(r)))
        return pid.value, a, r.value

    def set_mask(self, mask):
        m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
        _chk(self.ctx.L.lsd_frame_set_mask(self.ctx.p, self.p, _ptr(m)))

    def set_counters(self, tracked, mapped):
        _chk(self.ctx.L.lsd_frame_set_counters(self.ctx.p, self.p, int(tracked), int(mapped)))

    def counters(self):
        a, b = C.c_int(), C.c_int()
        _chk(self.ctx.L.lsd_frame_get_counters(self.ctx.p, self.p, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_depth_updated_flag(self, v):
        _chk(self.ctx.L.lsd_frame_set_depth_updated_flag(self.ctx.p, self.p, int(v)))

    def depth_updated_flag(self):
        v = C.c_int()
        _chk(self.ctx.L.lsd_frame_get_depth_updated_flag(self.ctx.p, self.p, C.byref(v)))
        return bool(v.value)

    def mean_idepth(self):
        m = C.c_float()
        n = C.c_int()
        _chk(self.ctx.L.lsd_frame_mean_idepth(self.ctx.p, self.p, C.byref(m), C.byref(n)))
        return m.value, n.value


class DepthMap:
    """[UP] lsd_slam::DepthMap: device-resident hypothesis planes; same method names as upstream."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        p = C.c_void_p()
        _chk(ctx.L.lsd_depthmap_create(ctx.p, C.byref(p)))
        self.p = p
        self._keep = []

    def destroy(self):
        if self.p:
            self.ctx.L.lsd_depthmap_destroy(self.ctx.p, self.p)
            self.p = None

    def set_thresholds(self, create=30, keep=24, unblacklist=100, min_blacklist=-1):
        s = DepthSettings(create, keep, unblacklist, min_blacklist)
        _chk(self.ctx.L.lsd_depthmap_set_settings(self.ctx.p, self.p, C.byref(s)))

    def initializeFromGTDepth(self, kf):
        self._keep.append(kf)
        _chk(self.ctx.L.lsd_depth_initialize_from_gt(self.ctx.p, self.p, kf.p))

    def initializeRandomly(self, kf):
        self._keep.append(kf)
        _chk(self.ctx.L.lsd_depth_initialize_randomly(self.ctx.p, self.p, kf.p))

    def initializeFromMap(self, kf, hyp, reactivated=False):
        self._keep.append(kf)
        h = np.ascontiguousarray(hyp, HYP_DTYPE)
        assert h.shape == (self.ctx.h, self.ctx.w)
        _chk(self.ctx.L.lsd_depth_initialize_from_map(self.ctx.p, self.p, kf.p, _ptr(h), int(reactivated)))

    def _frames(self, frames, refToKf):
        n = len(frames)
        fp = (C.c_void_p * n)(*[f.p for f in frames])
        poses = np.ascontiguousarray(refToKf, np.float64).reshape(n, 8) if refToKf is not None else None
        return n, fp, poses

    def updateKeyframe(self, referenceFrames, refToKf=None):
        n, fp, poses = self._frames(referenceFrames, refToKf)
        _chk(self.ctx.L.lsd_depth_update_keyframe(self.ctx.p, self.p, n, fp, _ptr(poses)))

    def prepare(self, referenceFrames, refToKf=None):
        n, fp, poses = self._frames(referenceFrames, refToKf)
        self._keep.extend(referenceFrames)
        _chk(self.ctx.L.lsd_depth_prepare(self.ctx.p, self.p, n, fp, _ptr(poses)))

    def createKeyFrame(self, new_keyframe):
        self._keep.append(new_keyframe)
        f = C.c_float()
        _chk(self.ctx.L.lsd_depth_create_keyframe(self.ctx.p, self.p, new_keyframe.p, C.byref(f)))
        return f.value

    def finalizeKeyFrame(self):
        _chk(self.ctx.L.lsd_depth_finalize_keyframe(self.ctx.p, self.p))

    def stage(self, stage, arg1=0, arg2=0, frame=None):
        if frame is not None:
            self._keep.append(frame)
        _chk(self.ctx.L.lsd_depth_stage(self.ctx.p, self.p, stage, arg1, arg2, frame.p if frame is not None else None))

    def read(self):
        out = np.zeros((self.ctx.h, self.ctx.w), HYP_DTYPE)
        _chk(self.ctx.L.lsd_depth_read(self.ctx.p, self.p, _ptr(out)))
        return out

    def debugPlotDepthMap(self):
        out = np.zeros((self.ctx.h, self.ctx.w, 3), np.uint8)
        _chk(self.ctx.L.lsd_depth_debug_rgb(self.ctx.p, self.p, _ptr(out)))
        return out


class Ref:
    """[UP] lsd_slam::TrackingReference."""

    def __init__(self, ctx: Context, p, kf: Frame):
        self.ctx, self.p, self.kf = ctx, p, kf

    def release(self):
        if self.p:
            self.ctx.L.lsd_ref_release(self.ctx.p, self.p)
            self.p = None

    def num_data(self, level):
        v = C.c_int()
        _chk(self.ctx.L.lsd_ref_num_data(self.ctx.p, self.p, level, C.byref(v)))
        return v.value

    def read(self, level):
        n = self.num_data(level)
        pos = np.empty((n, 3), np.float32)
        grad = np.empty((n, 2), np.float32)
        cv = np.empty((n, 2), np.float32)
        idx = np.empty((n,), np.int32)
        _chk(self.ctx.L.lsd_ref_read(self.ctx.p, self.p, level, _ptr(pos), _ptr(grad), _ptr(cv), _ptr(idx)))
        return pos, grad, cv, idx


class Undistorter:
    """libvideoio::Undistorter (OpenCV fixed-point remap) on the device; output size = the context's size."""

    def __init__(self, ctx: Context, in_w, in_h, maps=None, K=None, dist=None, K_out=None):
        self.ctx, self.in_w, self.in_h = ctx, in_w, in_h
        p = C.c_void_p()
        if maps is not None:
            m1 = np.ascontiguousarray(maps[0], np.int16)
            m2 = np.ascontiguousarray(maps[1], np.uint16)
            assert m1.shape == (ctx.h, ctx.w, 2) and m2.shape == (ctx.h, ctx.w)
            _chk(ctx.L.lsd_undistorter_create_from_maps(ctx.p, in_w, in_h, _ptr(m1), _ptr(m2), C.byref(p)))
        else:
            Kd = np.ascontiguousarray(K, np.float64)
            dd = np.zeros(5)
            dd[:len(dist)] = dist
            Ko = np.ascontiguousarray(K_out if K_out is not None else ctx.K, np.float64)
            _chk(ctx.L.lsd_undistorter_create_opencv(ctx.p, in_w, in_h, _ptr(Kd), _ptr(dd), _ptr(Ko), C.byref(p)))
        self.p = p

    def close(self):
        if self.p:
            self.ctx.L.lsd_undistorter_destroy(self.ctx.p, self.p)
            self.p = None

    def maps(self):
        m1 = np.zeros((self.ctx.h, self.ctx.w, 2), np.int16)
        m2 = np.zeros((self.ctx.h, self.ctx.w), np.uint16)
        _chk(self.ctx.L.lsd_undistorter_maps(self.ctx.p, self.p, _ptr(m1), _ptr(m2)))
        return m1, m2

    def undistort(self, image):
        im = np.ascontiguousarray(image, np.uint8)
        assert im.shape == (self.in_h, self.in_w)
        out = np.zeros((self.ctx.h, self.ctx.w), np.uint8)
        _chk(self.ctx.L.lsd_undistort(self.ctx.p, self.p, _ptr(im), self.in_w, _ptr(out)))
        return out

    def create_frames(self, images, ids=None, flags=BUILD_TRACKING, want_undistorted=False):
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        idarr = (C.c_int * n)(*ids) if ids is not None else None
        outs, up = None, None
        if want_undistorted:
            outs = [np.zeros((self.ctx.h, self.ctx.w), np.uint8) for _ in range(n)]
            up = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        fr = (C.c_void_p * n)()
        _chk(self.ctx.L.lsd_frame_create_undistorted_batch(self.ctx.p, self.p, n, idarr, ptrs, self.in_w, flags, up, fr))
        frames = [Frame(self.ctx, fr[i], ids[i] if ids is not None else i) for i in range(n)]
        return (frames, outs) if want_undistorted else frames


class SlamSystem:
    """[UP] lsd_slam::SlamSystem, lock-step (runRealTime == false): the native driver in csrc/slam.cu."""

    def __init__(self, ctx: Context, keep_keyframes=True):
        self.ctx = ctx
        p = C.c_void_p()
        _chk(ctx.L.lsd_slam_create(ctx.p, C.byref(p)))
        self.p = p
        _chk(ctx.L.lsd_slam_set_keep_keyframes(self.p, int(keep_keyframes)))
        self.lines = []

    def close(self):
        if self.p:
            self.ctx.L.lsd_slam_destroy(self.p)
            self.p = None

    def set_undistorter(self, und):
        _chk(self.ctx.L.lsd_slam_set_undistorter(self.p, und.p if und is not None else None))

    def set_pipelined(self, enable):
        _chk(self.ctx.L.lsd_slam_set_pipelined(self.p, int(bool(enable))))

    def _done(self, st):
        if st.tracked:
            buf = C.create_string_buffer(256)
            _chk(self.ctx.L.lsd_slam_pose_line(C.byref(st), buf, 256))
            self.lines.append(buf.value.decode().rstrip("\n"))
        return st

    def gtDepthInit(self, image, fid, depth):
        im = np.ascontiguousarray(image, np.uint8)
        d = np.ascontiguousarray(depth, np.float32)
        st = SlamStatus()
        _chk(self.ctx.L.lsd_slam_gt_depth_init(self.p, int(fid), _ptr(im), im.shape[1], _ptr(d), C.byref(st)))
        return self._done(st)

    def nextImage(self, image, fid):
        im = np.ascontiguousarray(image, np.uint8)
        st = SlamStatus()
        _chk(self.ctx.L.lsd_slam_next_image(self.p, int(fid), _ptr(im), im.shape[1], C.byref(st)))
        return self._done(st)

    @staticmethod
    def nextImageBatch(systems, images, fids, image_ptrs=None, pitch=None):
        """lsd_slam_next_image_batch: one image for each of n systems that share a context (every stage batched over the
        sequences).  images: n (h, w) uint8 arrays, or pass image_ptrs (host addresses) + pitch to skip the conversion."""
        n = len(systems)
        ctx = systems[0].ctx
        if image_ptrs is None:
            ims = [np.ascontiguousarray(im, np.uint8) for im in images]
            image_ptrs = [im.ctypes.data for im in ims]
            pitch = ims[0].shape[1]
        sp = (C.c_void_p * n)(*[s.p for s in systems])
        ip = (C.c_void_p * n)(*image_ptrs)
        idv = (C.c_int * n)(*[int(f) for f in fids])
        st = (SlamStatus * n)()
        _chk(ctx.L.lsd_slam_next_image_batch(n, sp, idv, ip, int(pitch), st))
        return [s._done(st[i]) for i, s in enumerate(systems)]

    def counters(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _chk(self.ctx.L.lsd_slam_counters(self.p, C.byref(a), C.byref(b), C.byref(c)))
        return dict(tracked=a.value, lost=b.value, keyframes=c.value - 1)

    def stage_seconds(self):
        out = np.zeros(5)
        _chk(self.ctx.L.lsd_slam_stage_seconds(self.p, _ptr(out)))
        return dict(zip(("ingest", "import_reference", "track", "update_keyframe", "keyframe_switch"), out.tolist()))

    def current_keyframe(self):
        kf = C.c_void_p()
        _chk(self.ctx.L.lsd_slam_current_keyframe(self.p, C.byref(kf), None))
        return Frame(self.ctx, kf, -1)
