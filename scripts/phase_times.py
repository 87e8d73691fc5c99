import sys, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from uasl_motion_estimation_b200 import capi, synth
for name in sys.argv[1:]:
    win = synth.config_window(name)
    cfg = capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=10)
    h = capi.Handle(cfg)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    h.time_iteration(2, iterations=3, flush_l2=False)
    h.set_profiling(True); h.timing(reset=True)
    ms = h.time_iteration(2, iterations=10, flush_l2=False)
    t = h.timing()
    print(name, "iter ms (profiled, serialised) %.3f"%ms, {k: round(v/10,4) for k,v in t.items() if k.endswith('_ms')})
    h.set_profiling(False)
    print(name, "iter ms (graph) %.3f"%h.time_iteration(2, iterations=10, flush_l2=False), "lin only %.4f"%h.time_linearize(2, 1e4, 10, False))
