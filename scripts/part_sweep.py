"""Lineariser time against the part size of k_lin_slot (UBA_SLOT_CAP=<chunks>[e]): calibrates the planner's cost model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
name = sys.argv[1]; caps = sys.argv[2:]
nwin = int(os.environ.get("SWEEP_WINDOWS", "512"))
wins = [synth.config_window("c3", window=i) for i in range(nwin)] if name == "c3" else [synth.config_window(name)]
batch = synth.concat_windows(wins) if len(wins) > 1 else None
nobs = sum(w.n_obs for w in wins)
for cap in ["model"] + caps:
    if cap == "model": os.environ.pop("UBA_SLOT_CAP", None)
    else: os.environ["UBA_SLOT_CAP"] = cap
    h = capi.Handle(capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=2))
    if batch is None:
        w = wins[0]; h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    else:
        h.set_batch(**batch)
    h.time_linearize(2, 1e4, 3, True)
    ms = min(h.time_linearize(2, 1e4, 20, True) for _ in range(3))
    print(name, "cap", cap, "lin ms %.4f" % ms, "iter ms %.4f" % h.time_iteration(2, 10, True), flush=True)
    h.close()
