"""Phase times of the cluster band solver (k_chol_banded_c2) from a -DUBA_BAND_TIMING build of libuba:
    nvcc ... -DUBA_BAND_TIMING -o /tmp/libuba_timing.so ...;  python scripts/band_timing.py <that .so> [c4|c5]"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
lib = capi.load(sys.argv[1])
lib.uba_debug_read_zbuf.argtypes = [C.c_void_p, capi.c_double_p, C.c_int]
win = synth.config_window(sys.argv[2] if len(sys.argv) > 2 else "c4", lib=lib)
cfg = capi.default_config(lib, fixed_iterations=3)
h = capi.Handle(cfg, lib=lib)
h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
h.optimise(2)
out = np.zeros(64)
lib.uba_debug_read_zbuf(h._h, capi.dptr(out), 64)
names = ["start", "loaded", "forward done", "cluster sync 1", "separator done", "cluster sync 2", "backward done"]
for half in range(2):
    v = out[half * 8:half * 8 + 7]
    print("CTA", half, {n: int(x) for n, x in zip(names, v)}, "deltas", [int(b - a) for a, b in zip(v[:-1], v[1:])])
print("CTA 0 busy cycles inside the forward loops (loop top -> step barrier), lane 0 of warps 0..7 (warp 7 = panel):", [int(x) for x in out[16:24]])
seg = ["prefetch issue", "row solves", "wait named barrier", "trailing update (DMMA)", "side duties", "reload store"]
for lbl, off in (("t0 (row solves)", 32), ("t100 (rhs update)", 40), ("t192 (rhs solve)", 48)):
    print(lbl, {n: int(v) for n, v in zip(seg, out[off:off + 6])})
print("CTA 0 backward: staging, block inverses (M, P^j), sweep [cycles]:", [int(x) for x in out[56:59]])
print("CTA 0 backward sweep, busy cycles before the step barrier: idle warp 0, helper warp, panel warp:", [int(x) for x in out[59:62]])
