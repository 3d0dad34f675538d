"""Per-frame comparison of the native (csrc/slam.cu) and Python lock-step drivers on the device (debug aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200")]
import numpy as np

import lsd_b200
from lsd_b200 import synth
from lsd_b200.pipeline import DeviceBackend, LockStepSlam

w, h = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (320, 240)))
order = sys.argv[3] if len(sys.argv) > 3 else "py_first"
K = synth.default_K(w, h)
room = synth.make_room(0)
traj = synth.trajectory(160, seed=0)[::4]
frames = [synth.render(room, w, h, K, R, t, noise_seed=i) for i, (R, t) in enumerate(traj)]
ctx = lsd_b200.Context(w, h, K)


def run_py():
    py = LockStepSlam(DeviceBackend(ctx))
    py.first_frame(frames[0][0].numpy(), 0, frames[0][1].numpy())
    out = []
    for i in range(1, len(frames)):
        k0 = py.stats["keyframes"]
        r = py.next_image(frames[i][0].numpy(), i)
        out.append((i, int(not (r.diverged or not r.trackingWasGood)), py.stats["keyframes"] - k0, r.pointUsage, r.lastResidual, r.lastGoodCount, r.lastBadCount))
    return out


def run_nat():
    nat = lsd_b200.SlamSystem(ctx)
    nat.gtDepthInit(frames[0][0].numpy(), 0, frames[0][1].numpy())
    out = []
    for i in range(1, len(frames)):
        s = nat.nextImage(frames[i][0].numpy(), i)
        out.append((i, s.tracked, s.isKeyframe, s.pointUsage, s.lastResidual, s.keyframeScore, s.diverged))
    nat.close()
    return out


a, b = (run_py(), run_nat()) if order == "py_first" else (None, run_nat())
if a is None:
    a = run_py()
for x, y in zip(a, b):
    flag = "" if (x[1], x[2]) == (y[1], y[2]) else "   <<<<"
    print("py", x[:3], "usage %.4f res %.4f good %.0f bad %.0f" % x[3:], "| nat", y[:3], "usage %.4f res %.4f score %.4f div %d" % y[3:], flag)

# ---- second native run on the SAME context: inspect the keyframe right after the first switch
nat = lsd_b200.SlamSystem(ctx)
nat.gtDepthInit(frames[0][0].numpy(), 0, frames[0][1].numpy())
for i in range(1, len(frames)):
    s = nat.nextImage(frames[i][0].numpy(), i)
    if s.isKeyframe or not s.tracked:
        kf = nat.current_keyframe()
        print("frame", i, "tracked", s.tracked, "isKF", s.isKeyframe, "kf id", s.currentKeyframeId)
        for l in range(5):
            idl, vl = kf.idepth(l), kf.idepthVar(l)
            print("  L%d valid %d nan %d mean %.4f" % (l, int((vl > 0).sum()), int(np.isnan(idl).sum()), float(idl[vl > 0].mean()) if (vl > 0).any() else -1))
        r = ctx.create_refs([kf])[0]
        print("  numData", [r.num_data(l) for l in (1, 2, 3, 4)], "meta", kf.tracking_meta()[1], kf.counters())
        if not s.tracked:
            break
nat.close()
ctx2 = lsd_b200.Context(w, h, K)
nat = lsd_b200.SlamSystem(ctx2)
nat.gtDepthInit(frames[0][0].numpy(), 0, frames[0][1].numpy())
lost = 0
for i in range(1, len(frames)):
    lost += 0 if nat.nextImage(frames[i][0].numpy(), i).tracked else 1
print("fresh context: lost", lost)
