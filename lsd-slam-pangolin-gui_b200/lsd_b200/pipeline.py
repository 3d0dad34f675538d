"""Lock-step tracking + mapping driver: the minimal part of [UP] lsd_slam::SlamSystem that
feeds the hot path -- SlamSystem::trackFrame, doMappingIteration (updateKeyframe / createNewCurrentKeyframe)
and the keyframe-selection score (A.10) -- with the reference's `--no-realtime` semantics (nextImage blocks
until the frame is mapped: /root/reference/lib/App/InputThread.cpp:70-71).  Pose-graph optimisation, loop
closure and relocalisation are NOT here (out of scope; they stay on the reference's CPU code).

The driver is written against a small backend protocol so that the SAME loop runs on the device
(DeviceBackend below) and, in tests / bench.py, on the CPU oracle; it emits the lines the reference writes to
pose.txt (`id,tx,ty,tz,rawtx,rawty,rawtz`, lib/Pangolin_IOWrapper/TextOutputIOWrapper.cpp:100-120).
"""
from __future__ import annotations

import numpy as np

# util/settings.h
KF_DIST_WEIGHT = 4.0
KF_USAGE_WEIGHT = 3.0
INITIALIZATION_PHASE_COUNT = 5
MIN_NUM_MAPPED = 5


def ref_frame_score(distance_squared, usage):
    """[UP] TrackableKeyFrameSearch::getRefFrameScore: distSq * KFDistWeight^2 + (1 - usage)^2 * KFUsageWeight^2, in float32."""
    f = np.float32
    d2, u = f(distance_squared), f(usage)
    w1, w2 = f(KF_DIST_WEIGHT), f(KF_USAGE_WEIGHT)
    return float(d2 * w1 * w1 + (f(1) - u) * (f(1) - u) * w2 * w2)


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_rot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v, np.float64)


def sim3_mul(a, b):
    """(q, t, s) composition of Sim3 double[8] {qx,qy,qz,qw,tx,ty,tz,s}: a * b"""
    q = quat_mul(a[:4], b[:4])
    q /= np.linalg.norm(q)
    t = a[4:7] + a[7] * quat_rot(a[:4], b[4:7])
    return np.concatenate([q, t, [a[7] * b[7]]])


def sim3_inv(a):
    qc = np.array([-a[0], -a[1], -a[2], a[3]])
    si = 1.0 / a[7]
    return np.concatenate([qc, -si * quat_rot(qc, a[4:7]), [si]])


IDENTITY8 = np.array([0, 0, 0, 1, 0, 0, 0, 1.0])


class DeviceBackend:
    """The product backend: lsd_b200.Context (CUDA kernels through the C ABI)."""

    def __init__(self, ctx):
        import lsd_b200
        self.ctx = ctx
        self.flags = lsd_b200.BUILD_MAXGRAD0

    def new_frame(self, img, fid):
        return self.ctx.create_frame(img, fid, flags=self.flags)

    def release_frame(self, f):
        f.release()

    def set_depth_gt(self, f, depth):
        f.set_depth_from_gt(depth)

    def new_depthmap(self):
        return self.ctx.create_depthmap()

    def init_gt(self, dm, kf):
        dm.initializeFromGTDepth(kf)

    def update_keyframe(self, dm, frames):
        dm.updateKeyframe(frames)

    def create_keyframe(self, dm, f):
        return dm.createKeyFrame(f)

    def finalize(self, dm):
        dm.finalizeKeyFrame()

    def import_ref(self, old, kf):
        if old is not None:
            old.release()
        return self.ctx.create_refs([kf])[0]

    def track(self, ref, f, init7):
        return self.ctx.se3_track(ref, f, init7)

    def depth_flag(self, kf):
        return kf.depth_updated_flag()

    def clear_depth_flag(self, kf):
        kf.set_depth_updated_flag(0)

    def mean_idepth(self, kf):
        return kf.mean_idepth()[0]

    def num_mapped(self, kf):
        return kf.counters()[1]

    def to_parent(self, f):
        return f.tracking_meta()[1]


class LockStepSlam:
    def __init__(self, backend):
        self.b = backend
        self.dm = backend.new_depthmap()
        self.kf = None
        self.ref = None
        self.ref_kf_id = None
        self.kf_world = IDENTITY8.copy()   # camToWorld of the current keyframe (Sim3)
        self.last_to_kf = IDENTITY8.copy()  # last tracked frame -> current keyframe
        self.n_keyframes = 0
        self.lines = []       # pose.txt lines
        self.world_poses = []  # (id, camToWorld[8])
        self.keyframe_ids = []
        self.stats = dict(tracked=0, lost=0, keyframes=0)
        self._frames = []

    def first_frame(self, img, fid, gt_depth):
        """SlamSystem::gtDepthInit: keyframe 0 with ground-truth depth (avoids rand())."""
        kf = self.b.new_frame(img, fid)
        self.b.set_depth_gt(kf, gt_depth)
        self.b.init_gt(self.dm, kf)
        self.kf = kf
        self.n_keyframes = 1
        self.keyframe_ids.append(fid)
        self._emit(fid, IDENTITY8, IDENTITY8)

    def _emit(self, fid, cam_to_world, raw):
        self.world_poses.append((fid, cam_to_world.copy()))
        t, r = cam_to_world[4:7], raw[4:7]
        self.lines.append(f"{fid},{t[0]:g},{t[1]:g},{t[2]:g},{r[0]:g},{r[1]:g},{r[2]:g}")

    def next_image(self, img, fid):
        """SlamSystem::trackFrame + one blocking doMappingIteration."""
        b = self.b
        f = b.new_frame(img, fid)
        # TrackingReference::importFrame when the keyframe changed or its depth was updated
        if self.ref is None or self.ref_kf_id != self.kf.id or b.depth_flag(self.kf):
            self.ref = b.import_ref(self.ref, self.kf)
            self.ref_kf_id = self.kf.id
            b.clear_depth_flag(self.kf)
        init = self.last_to_kf[:7].copy()  # frameToReference_initialEstimate: last tracked pose relative to the keyframe
        res = b.track(self.ref, f, init)
        if res.diverged or not res.trackingWasGood:
            self.stats["lost"] += 1  # upstream would start the Relocalizer (out of scope): keep the keyframe, drop the frame
            b.release_frame(f)
            return res
        self.stats["tracked"] += 1
        to_kf = np.concatenate([np.array(res.frameToRef), [1.0]])
        self.last_to_kf = to_kf
        self._emit(fid, sim3_mul(self.kf_world, to_kf), to_kf)

        # keyframe selection (A.10)
        create = False
        if b.num_mapped(self.kf) > MIN_NUM_MAPPED:
            dist = to_kf[4:7] * b.mean_idepth(self.kf)
            min_val = min(0.2 + self.n_keyframes * 0.8 / INITIALIZATION_PHASE_COUNT, 1.0)
            if self.n_keyframes < INITIALIZATION_PHASE_COUNT:
                min_val *= 0.7
            score = ref_frame_score(float(dist @ dist), res.pointUsage)
            create = score > min_val
        # mapping, lock-step
        if create:
            b.finalize(self.dm)          # finishCurrentKeyframe
            b.create_keyframe(self.dm, f)  # createNewCurrentKeyframe
            raw = np.array(b.to_parent(f))  # now carries the rescale factor
            self.kf_world = sim3_mul(self.kf_world, raw)
            old = self.kf
            self.kf = f
            self.n_keyframes += 1
            self.keyframe_ids.append(fid)
            self.last_to_kf = IDENTITY8.copy()
            self._frames.append(old)  # keyframes stay alive (upstream keeps them in the graph)
            self.stats["keyframes"] += 1
        else:
            b.update_keyframe(self.dm, [f])
            b.release_frame(f)
        return res
