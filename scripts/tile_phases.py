import sys
sys.path.insert(0,'/root/repo')
from uasl_motion_estimation_b200 import capi, synth
for name in ("c4", "c1"):
    for m in (1, 3, 7, 15, 31, 255):
        lib = capi.load(f'scripts/libuba_ph{m}.so')
        wins = [synth.config_window(name, lib=lib)] if name == "c4" else [synth.config_window("c3", window=i, lib=lib) for i in range(148)]
        h = capi.Handle(capi.default_config(lib, fixed_iterations=2), lib=lib)
        if len(wins) == 1:
            w = wins[0]; h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
        else:
            h.set_batch(**synth.concat_windows(wins))
        h.time_linearize(2, 1e4, 3, False)
        print(name if name == "c4" else "c3x148", "phases mask", m, "lin ms %.4f" % h.time_linearize(2, 1e4, 10, False), flush=True)
