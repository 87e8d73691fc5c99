import sys
sys.path.insert(0,'/root/repo')
from uasl_motion_estimation_b200 import capi, synth
for name in sys.argv[1:]:
    if name == "c3":
        wins = [synth.config_window("c3", window=i) for i in range(296)]
    else:
        wins = [synth.config_window(name)]
    h = capi.Handle(capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=2))
    if len(wins) == 1:
        w = wins[0]; h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    else:
        h.set_batch(**synth.concat_windows(wins))
    h.time_linearize(2, 1e4, 3, False)
    nobs = sum(w.n_obs for w in wins)
    ms = h.time_linearize(2, 1e4, 20, False)
    print(name, "lin ms %.4f  obs/s %.3e" % (ms, nobs / ms * 1e3), "iter ms %.4f" % h.time_iteration(2, 10, False), flush=True)
