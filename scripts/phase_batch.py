"""Per-phase CUDA-event times of one LM iteration for every BASELINE workload (c3: 512 windows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, synth
for name in sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]:
    wins = [synth.config_window("c3", window=i) for i in range(512)] if name == "c3" else [synth.config_window(name)]
    h = capi.Handle(capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=10))
    if len(wins) == 1:
        w = wins[0]; h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    else:
        h.set_batch(**synth.concat_windows(wins))
    h.time_iteration(2, iterations=3, flush_l2=False)
    h.set_profiling(True); h.timing(reset=True)
    ms = h.time_iteration(2, iterations=10, flush_l2=False)
    t = h.timing()
    h.set_profiling(False)
    print(name, "serialised %.3f" % ms, {k: round(v / 10, 4) for k, v in t.items() if k.endswith('_ms') and v}, "graph %.3f" % h.time_iteration(2, iterations=10, flush_l2=True), flush=True)
