"""The CPU-oracle backend of lsd_b200.pipeline.LockStepSlam (test infrastructure: tests/ and bench.py only)."""
import numpy as np

from oracle import pyoracle as O


class OracleBackend:
    def __init__(self, w, h, K, mode=2, threads=1, fast=False):
        self.w, self.h, self.K, self.mode, self.threads, self.fast = w, h, K, mode, threads, fast

    def new_frame(self, img, fid):
        f = O.Frame(fid, img, self.K, fast=self.fast)
        f.build_pyramids()
        return f

    def release_frame(self, f):
        pass

    def set_depth_gt(self, f, depth):
        f.set_depth_gt(depth)

    def new_depthmap(self):
        return O.DepthMap(self.w, self.h, self.K, threads=self.threads, fast=self.fast)

    def init_gt(self, dm, kf):
        dm.init_gt(kf)

    def update_keyframe(self, dm, frames):
        dm.update_keyframe(frames)

    def create_keyframe(self, dm, f):
        dm.create_keyframe(f)
        return dm.last_rescale()

    def finalize(self, dm):
        dm.finalize()

    def import_ref(self, old, kf):
        return O.Ref(kf)

    def track(self, ref, f, init7):
        res, _ = O.se3_track(ref, f, init7, self.mode)
        return res

    def depth_flag(self, kf):
        return bool(kf.L.lsdo_frame_get_flags(kf.p))

    def clear_depth_flag(self, kf):
        kf.L.lsdo_frame_set_flags(kf.p, 0)

    def mean_idepth(self, kf):
        return kf.mean_idepth()

    def num_mapped(self, kf):
        return int(O.frame_counters(kf)[1])

    def to_parent(self, f):
        return O.frame_pose(f)
