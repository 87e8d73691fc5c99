"""Host-side breakdown of one submit + uba_optimise + read-back call (UBA_TRACE=1 prints the ingest phases on stderr).
c3 goes through uba_set_batch with the concatenated arrays prepared once, as bench.py's e2e leg does."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UBA_TRACE"] = "1"
import numpy as np
import bench
from uasl_motion_estimation_b200 import capi, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wins, total_obs, scaling, parallelism, n_local = bench.build_workload(name, 0, 1, 0, 1.0)
cfg = capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=synth.CONFIGS[name]["iters"])
h = capi.Handle(cfg)
batch = synth.concat_windows(wins) if len(wins) > 1 else None
cams_out = pts_out = None
for rep in range(5):
    print(f"--- {name} rep {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    if batch is None:
        w = wins[0]
        h.set_problem(w.M, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    else:
        h.set_batch(**batch)
    t1 = time.perf_counter()
    rc, sums = h.optimise(2)
    t2 = time.perf_counter()
    if cams_out is None:
        cams_out = np.zeros((h.n_cams, 6)); pts_out = np.zeros((h.n_pts, 3))
    c = h.cameras(cams_out); p = h.points(pts_out)
    t3 = time.perf_counter()
    print(f"rep {rep}: submit {1e3*(t1-t0):.2f} ms, optimise {1e3*(t2-t1):.2f} ms, read-back {1e3*(t3-t2):.2f} ms, total {1e3*(t3-t0):.2f} ms", file=sys.stderr, flush=True)
