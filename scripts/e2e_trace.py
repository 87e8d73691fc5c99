"""Host-side breakdown of one uba_set_problem + uba_optimise + read-back call (UBA_TRACE=1 prints the ingest phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UBA_TRACE"] = "1"
import numpy as np
from uasl_motion_estimation_b200 import capi, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
win = synth.config_window(name)
cfg = capi.default_config(fixed_iterations=10)
h = capi.Handle(cfg)
for rep in range(4):
    t0 = time.perf_counter()
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    t1 = time.perf_counter()
    rc, sums = h.optimise(2)
    t2 = time.perf_counter()
    c = h.cameras(); p = h.points()
    t3 = time.perf_counter()
    print(f"rep {rep}: set_problem {1e3*(t1-t0):.2f} ms, optimise {1e3*(t2-t1):.2f} ms, read-back {1e3*(t3-t2):.2f} ms, total {1e3*(t3-t0):.2f} ms", file=sys.stderr)
