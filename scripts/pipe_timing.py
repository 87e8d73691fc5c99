"""When does the band solver run relative to the lineariser in the pipelined iteration?  globaltimer marks from a
-DUBA_BAND_TIMING build:  python scripts/pipe_timing.py <that .so> [c4|c5]"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
lib = capi.load(sys.argv[1])
lib.uba_debug_read_zbuf.argtypes = [C.c_void_p, capi.c_double_p, C.c_int]
win = synth.config_window(sys.argv[2] if len(sys.argv) > 2 else "c4", lib=lib)
h = capi.Handle(capi.default_config(lib, fixed_iterations=3), lib=lib)
h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
h.optimise(2)
out = np.zeros(2100)
lib.uba_debug_read_zbuf(h._h, capi.dptr(out), 2100)
names = ["start", "loaded", "forward done", "cluster sync 1", "separator done", "cluster sync 2", "backward done"]
st = out[100:1100]; en = out[1100:2100]
n = int((en > 0).sum())
t0 = st[:n].min()
print(f"lineariser (last iteration): {n} CTAs, first start 0, last start {st[:n].max() - t0:.0f} ns, first end {en[:n].min() - t0:.0f}, last end {en[:n].max() - t0:.0f} ns")
q = np.sort(en[:n] - t0); print("   CTA end times, deciles [us]:", [round(float(q[int(i * (n - 1) / 10)]) / 1e3, 1) for i in range(11)])
for half in range(2):
    g = out[64 + half * 8: 64 + half * 8 + 7]
    print("solver CTA", half, {nm: round((x - t0) / 1e3, 1) for nm, x in zip(names, g)}, "[us after the first lineariser CTA started]")
