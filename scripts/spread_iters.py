"""Run-to-run spread of poses and points against the LM iteration count (floating-point atomics reorder sums; weakly constrained
points amplify that noise).  python scripts/spread_iters.py c5"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uasl_motion_estimation_b200 import capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
w = synth.config_window(name)
for iters in (1, 4, 12, 20):
    cfg = capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=iters)
    h = capi.Handle(cfg)
    ref = None; wc = 0.0; dp = None
    for r in range(8):
        h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
        rc, s = h.optimise(2)
        c, p = h.cameras(), h.points()
        if ref is None: ref = (c.copy(), p.copy()); dp = np.zeros(len(p))
        else:
            wc = max(wc, np.abs(c - ref[0]).max() / np.abs(ref[0]).max())
            dp = np.maximum(dp, np.abs(p - ref[1]).max(axis=1) / np.abs(ref[1]).max())
    worst = int(np.argmax(dp)); nobs = int((w.pt_idx == worst).sum())
    print(f"{name} iters {iters:2d}: poses {wc:.1e}; points: max {dp.max():.1e}, median {np.median(dp):.1e}, above 1e-10: {(dp > 1e-10).sum()} of {len(dp)}"
          f"; worst point {worst} has {nobs} observations, |X| = {np.abs(ref[1][worst]).max():.1f}", flush=True)
