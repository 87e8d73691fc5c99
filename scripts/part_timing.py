"""Per-CTA timeline of k_lin_slot from a -DUBA_BAND_TIMING build (globaltimer marks: start, first chunk's data there, flush
start, end):  python scripts/part_timing.py <that .so> [c4|c5]   (UBA_SLOT_CAP etc. apply)"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
lib = capi.load(sys.argv[1])
lib.uba_debug_read_zbuf.argtypes = [C.c_void_p, capi.c_double_p, C.c_int]
name = sys.argv[2] if len(sys.argv) > 2 else "c4"
win = synth.config_window(name, lib=lib)
h = capi.Handle(capi.default_config(lib, loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=3), lib=lib)
h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
h.optimise(2)
out = np.zeros(4100)
lib.uba_debug_read_zbuf(h._h, capi.dptr(out), 4100)
st, en, data, fl = out[100:1100], out[1100:2100], out[2100:3100], out[3100:4100]
n = int((en > 0).sum())
t0 = st[:n].min()
us = lambda x: (x[:n] - t0) / 1e3
S, E, D, F = us(st), us(en), us(data), us(fl)
print(f"{n} CTAs; kernel span {E.max():.1f} us")
print("  start -> first data [us]: median %.2f  p90 %.2f" % (np.median(D - S), np.percentile(D - S, 90)))
print("  flush [us]:               median %.2f  p90 %.2f" % (np.median(E - F), np.percentile(E - F, 90)))
print("  whole part [us]:          median %.2f  p10 %.2f  p90 %.2f  max %.2f" % (np.median(E - S), np.percentile(E - S, 10), np.percentile(E - S, 90), (E - S).max()))
order = np.argsort(S)
print("  start times deciles:", [round(float(np.sort(S)[int(i * (n - 1) / 10)]), 1) for i in range(11)])
print("  end times deciles:  ", [round(float(np.sort(E)[int(i * (n - 1) / 10)]), 1) for i in range(11)])
first = S < 2.0
print("  first wave: %d CTAs, duration median %.1f; later CTAs: %d, duration median %.1f" % (first.sum(), np.median((E - S)[first]), (~first).sum(), np.median((E - S)[~first]) if (~first).any() else 0))
for b in range(0, n, max(1, n // 24)):
    print("   cta %4d start %6.1f data %6.1f flush %6.1f end %6.1f" % (b, S[b], D[b], F[b], E[b]))
