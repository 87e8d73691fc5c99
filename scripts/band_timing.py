import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from uasl_motion_estimation_b200 import capi, synth
lib = capi.load(sys.argv[1])
lib.uba_debug_read_zbuf.argtypes=[C.c_void_p, capi.c_double_p, C.c_int]
win = synth.config_window("c4", lib=lib)
cfg = capi.default_config(lib, fixed_iterations=3)
h = capi.Handle(cfg, lib=lib)
h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
h.optimise(2)
out = np.zeros(40)
lib.uba_debug_read_zbuf(h._h, capi.dptr(out), 40)
names = ["panel_T6","panel_corner","panel_factor","w_trsm","w_namedbar","w_update","cta_barrier","-"]
for lbl, off in (("t0 (row thread)",0),("t64 (y/Lt thread)",8),("t100 (pairs only)",16),("t224 (panel lane 0)",24)):
    print(lbl, {n:int(v) for n,v in zip(names, out[off:off+8])}, "sum", int(out[off:off+8].sum()))
print("backward cycles", int(out[32]))
