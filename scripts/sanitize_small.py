"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): c1 and c2 shapes (dense small solver), a 60-keyframe banded
window (cluster band solver, tiled lineariser), a batch, and the covariance pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uasl_motion_estimation_b200 import capi, synth

def run(win, **kw):
    cfg = capi.default_config(fixed_iterations=2, **kw)
    h = capi.Handle(cfg)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    rc, s = h.optimise(2)
    assert rc == 0, rc
    return h.cameras(), s[0].final_cost

print("c1", run(synth.config_window("c1", scale=0.05))[1])
print("c2", run(synth.config_window("c2", scale=0.02), compute_covariance=1)[1])
big = synth.generate(60, 1500, 5, 5, seed=7)
print("banded 60 keyframes", run(big)[1])
print("banded, generic lineariser", run(big, linearizer=1)[1])
wins = [synth.config_window("c1", window=i, scale=0.03) for i in range(3)]
h = capi.Handle(capi.default_config(fixed_iterations=2)); h.set_batch(**synth.concat_windows(wins)); print("batch", h.optimise(2)[0])
print("SANITIZE RUN OK")
