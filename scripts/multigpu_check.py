"""torchrun --nproc-per-node N scripts/multigpu_check.py : point-sharded c4 over NCCL vs the single-GPU run."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uasl_motion_estimation_b200 import capi, sharding, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
win = synth.config_window("c4", scale=scale)
cfg = capi.default_config(fixed_iterations=6, device=local)
h = capi.Handle(cfg)
uid = [h.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
h.comm_init(uid[0], rank, world)
sh = sharding.shard_window(win, rank, world)
h.set_problem(4, sh.cams_init, sh.pts_init, sh.feats, sh.cam_idx, sh.pt_idx, sh.cam_id, sh.calib)
rc, sums = h.optimise(2)
cams = h.cameras(); pts = h.points()
b = sharding.point_ranges(win.pt_idx, win.n_pts, world)
if True:
    h1 = capi.Handle(capi.default_config(fixed_iterations=6, device=local))
    h1.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    rc1, sums1 = h1.optimise(2)
    c1 = h1.cameras(); p1 = h1.points()[b[rank]:b[rank + 1]]
    ec = np.abs(cams - c1).max() / np.abs(c1).max(); ep = np.abs(pts - p1).max() / np.abs(p1).max()
    acc = [it["accepted"] for it in h.iterations(0)]; acc1 = [it["accepted"] for it in h1.iterations(0)]
    print(f"rank {rank}/{world}: rc {rc}/{rc1} cost {sums[0].final_cost:.9e} vs {sums1[0].final_cost:.9e} cams rel {ec:.2e} pts rel {ep:.2e} accepted {acc} {acc1}", flush=True)
    assert rc == 0 and ec < 1e-6 and ep < 1e-6 and acc == acc1
dist.barrier()
if rank == 0:
    print("MULTIGPU CHECK OK")
dist.destroy_process_group()
