"""Clock marks of the block-cyclic-reduction band solver (k_chol_bcr) from a -DUBA_BAND_TIMING build of libuba:
    python scripts/bcr_timing.py <that .so> [c4|c5]     (marks: start, phase 0, then before / after every cluster barrier)"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
lib = capi.load(sys.argv[1])
lib.uba_debug_read_zbuf.argtypes = [C.c_void_p, capi.c_double_p, C.c_int]
win = synth.config_window(sys.argv[2] if len(sys.argv) > 2 else "c4", lib=lib)
h = capi.Handle(capi.default_config(lib, fixed_iterations=3), lib=lib)
h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
h.optimise(2)
out = np.zeros(64)
lib.uba_debug_read_zbuf(h._h, capi.dptr(out), 64)
for r in range(2):
    v = out[r * 24:r * 24 + 24]
    print("CTA", r, "marks", [int(x) for x in v if x > 0 or x is v[0]])
    print("   deltas", [int(b - a) for a, b in zip(v[:-1], v[1:]) if b > 0])
