import sys
sys.path.insert(0,'/root/repo')
from uasl_motion_estimation_b200 import capi, synth
lib = capi.load(sys.argv[1])
for name in sys.argv[2:]:
    win = synth.config_window(name, lib=lib)
    cfg = capi.default_config(lib, loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=10)
    h = capi.Handle(cfg, lib=lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    h.time_iteration(2, iterations=3, flush_l2=False)
    h.set_profiling(True); h.timing(reset=True)
    ms = h.time_iteration(2, iterations=10, flush_l2=False)
    t = h.timing()
    print(sys.argv[1].split('/')[-1], name, {k: round(v/10,4) for k,v in t.items() if k.endswith('_ms') and v})
