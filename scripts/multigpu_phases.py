"""torchrun --nproc-per-node N scripts/multigpu_phases.py [workload]: per-phase CUDA-event times of a point-sharded window on
every rank (profiling mode: phases serialised; comm = the two peer-memory exchanges, which include waiting for the slowest
rank), next to the graph-replayed iteration time."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ["WORLD_SIZE"])))
from uasl_motion_estimation_b200 import capi, sharding, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
win = synth.config_window(name)
sh = sharding.shard_window(win, rank, world)
h = capi.Handle(capi.default_config(device=local, loss_kind=synth.CONFIGS[name]["loss"], fixed_iterations=10))
uid = [h.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
h.comm_init(uid[0], rank, world)
h.set_problem(4, sh.cams_init, sh.pts_init, sh.feats, sh.cam_idx, sh.pt_idx, sh.cam_id, sh.calib)
h.time_iteration(2, iterations=3, flush_l2=True)
for flush in (True, False):
    dist.barrier(); torch.cuda.synchronize()
    g = h.time_iteration(2, iterations=20, flush_l2=flush)
    print(f"rank {rank}/{world}: graph iteration {g:.4f} ms (L2 flush {flush}), {sh.n_obs} obs {sh.n_pts} pts", flush=True)
dist.barrier(); torch.cuda.synchronize()
h.set_profiling(True); h.timing(reset=True)
ms = h.time_iteration(2, iterations=10, flush_l2=False)
t = h.timing()
h.set_profiling(False)
print(f"rank {rank}/{world}: serialised {ms:.4f}", {k: round(v / 10, 4) for k, v in t.items() if k.endswith('_ms') and v}, flush=True)
dist.barrier()
