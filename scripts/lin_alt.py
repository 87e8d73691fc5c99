import sys
sys.path.insert(0,'/root/repo')
from uasl_motion_estimation_b200 import capi, synth
capi._default_lib = capi.load(sys.argv[1])
for name in sys.argv[2:]:
    w = synth.config_window(name)
    h = capi.Handle(capi.default_config(loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=2))
    h.set_problem(4, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    h.time_linearize(2, 1e4, 3, False)
    print(sys.argv[1].split('/')[-1], name, "lin ms warm %.4f cold %.4f" % (h.time_linearize(2, 1e4, 20, False), h.time_linearize(2, 1e4, 20, True)), flush=True)
