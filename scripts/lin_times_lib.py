import sys
sys.path.insert(0,'/root/repo')
from uasl_motion_estimation_b200 import capi, synth
libpath = sys.argv[1]
lib = capi.load(libpath)
for name in sys.argv[2:]:
    wins = [synth.config_window("c3", window=i, lib=lib) for i in range(296)] if name == "c3" else [synth.config_window(name, lib=lib)]
    h = capi.Handle(capi.default_config(lib, loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=2), lib=lib)
    if len(wins) == 1:
        w = wins[0]; h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    else:
        h.set_batch(**synth.concat_windows(wins))
    h.time_linearize(2, 1e4, 3, False)
    nobs = sum(w.n_obs for w in wins)
    ms = h.time_linearize(2, 1e4, 20, False)
    print(libpath.split('/')[-1], name, "lin ms %.4f  obs/s %.3e" % (ms, nobs / ms * 1e3), flush=True)
