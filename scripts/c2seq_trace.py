"""Host-side phase times (UBA_TRACE=1) of a few per-frame calls of the c2 sequence: sliding (uba_window_advance) and re-submitted."""
import os, sys, time
os.environ["UBA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uasl_motion_estimation_b200 import capi, synth
seq = synth.SlidingSequence()
for sliding in (1, 0):
    ids = seq.initial_ids(); w = seq.window(0, ids)
    h = capi.Handle(capi.default_config(fixed_iterations=4, sliding_window=sliding))
    h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib); h.optimise(2)
    cams, pts = h.cameras(), h.points()
    for first in range(60):
        kw, ids_new = seq.advance(first, ids)
        trace = first >= 57
        if not sliding:
            alive = seq.hi[ids] >= first + 1
            w = seq.window(first + 1, ids_new, cams=np.concatenate([cams[1:], kw["new_cams6"]]), pts=np.concatenate([pts[alive], kw["new_pts3"]]))
        if not trace:
            fd = os.dup(2); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 2)
        t0 = time.perf_counter()
        if sliding: h.window_advance(**kw)
        else: h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
        t1 = time.perf_counter(); h.optimise(2); t2 = time.perf_counter(); cams, pts = h.cameras(), h.points(); t3 = time.perf_counter()
        if not trace:
            os.dup2(fd, 2); os.close(fd); os.close(dn)
        else:
            print(f"== {'sliding' if sliding else 'resubmit'} call {first}: submit {1e3*(t1-t0):.2f} optimise {1e3*(t2-t1):.2f} read-back {1e3*(t3-t2):.2f} ms, {h.n_obs} obs {h.n_pts} pts", flush=True)
        ids = ids_new
