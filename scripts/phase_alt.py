"""phase times with an alternative build of the library: python scripts/phase_alt.py <lib.so> c4 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
capi._default_lib = capi.load(sys.argv[1])
for name in sys.argv[2:]:
    w = synth.config_window(name)
    h = capi.Handle(capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=12))
    h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    h.optimise(2)
    h.set_profiling(True); h.timing(reset=True)
    h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    h.optimise(2)
    t = h.timing()
    print(os.path.basename(sys.argv[1]), name, {k: round(v / 12, 4) for k, v in t.items() if k.endswith("_ms") and v}, flush=True)
