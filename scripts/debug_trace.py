import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lsd-slam-pangolin-gui_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import lsd_b200
from common import make_oracle_pair
from oracle import pyoracle as O
seed, w, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = make_oracle_pair(seed, w, h)
ctx = lsd_b200.Context(w, h, d["pr"]["K"])
kf = ctx.create_frame(d["kf_img"], 0); fr = ctx.create_frame(d["fr_img"], 1)
kf.set_idepth(d["idepth"], d["var"]); ref = ctx.create_refs([kf])[0]
init = np.array([0, 0, 0, 1, 0, 0, 0.0])
g, gt = ctx.se3_track(ref, fr, init, want_trace=True)
e, et = O.se3_track(d["oref"], d["ofr"], init, 2)
o, ot = O.se3_track(d["oref"], d["ofr"], init, 0)
for k, (a, b, c) in enumerate(zip(gt, et, ot)):
    print(k, a[0], a[1], b[1], c[1], 'size', a[4], b[4], 'err %.8g %.8g %.8g' % (a[2], b[2], c[2]), 'rel g-e %.2e  o-e %.2e' % ((a[2]-b[2])/b[2], (c[2]-b[2])/b[2]), 'lam', a[3], b[3])
print('gpu', np.array(g.frameToRef)); print('exa', np.array(e.frameToRef)); print('sca', np.array(o.frameToRef))
print('aff', g.affine_a, g.affine_b, e.affine_a, e.affine_b)
