#!/usr/bin/env python
"""Small lock-step run for compute-sanitizer (scripts: the GPU tests render their sequences on the device, which is too slow
under the sanitizer): 320x240, frames rendered on the CPU, the pipelined single-sequence driver and the batched driver."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"))
import lsd_b200  # noqa: E402
from lsd_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
w, h = 320, 240
K = synth.default_K(w, h)
room = synth.make_room(0)
traj = synth.trajectory(4 * n, seed=0)[::4]
frames = []
for i, (R, t) in enumerate(traj):
    img, depth = synth.render(room, w, h, K, R, t, noise_seed=i)
    frames.append((np.ascontiguousarray(img.cpu().numpy()), depth.cpu().numpy()))
ctx = lsd_b200.Context(w, h, K)
ctx.set_live_tracking(True)
s = lsd_b200.SlamSystem(ctx)
s.gtDepthInit(frames[0][0], 0, frames[0][1])
for i in range(1, n):
    st = s.nextImage(frames[i][0], i)
kf = s.current_keyframe()
print("single:", s.counters(), float(np.nanmax(kf.idepth(0))))
s.close()
sys2 = [lsd_b200.SlamSystem(ctx) for _ in range(3)]
for q in sys2:
    q.gtDepthInit(frames[0][0], 0, frames[0][1])
for i in range(1, n):
    lsd_b200.SlamSystem.nextImageBatch(sys2, [frames[i][0]] * 3, [i] * 3)
print("batch:", [q.counters() for q in sys2])
for q in sys2:
    q.close()
ctx.close()
print("done")
