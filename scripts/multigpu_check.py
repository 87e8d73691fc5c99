"""torchrun --nproc-per-node N scripts/multigpu_check.py [scale] [workload] : a point-sharded window over N GPUs against
the single-GPU run of the same window and against the CPU oracle (tests/test_multigpu.py runs this under pytest).

Checks, on every rank: (1) fixed-K trajectory: accept/reject sequence, poses and this rank's points equal to the single-GPU
run and to the oracle within 1e-6; (2) the reference's own termination rules (no fixed K, wall-clock cap off): same iteration
count and termination on every rank; (3) pose covariances on a sharded handle equal to the single-GPU ones."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from uasl_motion_estimation_b200 import capi, sharding, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
name = sys.argv[2] if len(sys.argv) > 2 else "c4"
K = synth.CONFIGS[name]["iters"]
loss = synth.CONFIGS[name]["loss"]
win = synth.config_window(name, scale=scale)
sh, ids = sharding.shard_window(win, rank, world, return_ids=True)
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())


def sharded_handle(**kw):
    h = capi.Handle(capi.default_config(device=local, loss_kind=loss, **kw))
    uid = [h.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    h.comm_init(uid[0], rank, world)
    h.set_problem(4, sh.cams_init, sh.pts_init, sh.feats, sh.cam_idx, sh.pt_idx, sh.cam_id, sh.calib)
    return h


def single_handle(**kw):
    h = capi.Handle(capi.default_config(device=local, loss_kind=loss, **kw))
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    return h


# (1) fixed-K trajectory
h = sharded_handle(fixed_iterations=K)
rc, sums = h.optimise(2)
cams = h.cameras(); pts = h.points()
h1 = single_handle(fixed_iterations=K)
rc1, sums1 = h1.optimise(2)
c1 = h1.cameras(); p1 = h1.points()[ids]
acc = [it["accepted"] for it in h.iterations(0)]; acc1 = [it["accepted"] for it in h1.iterations(0)]
msg = f"rank {rank}/{world}: K={K} rc {rc}/{rc1} cost {sums[0].final_cost:.9e} vs {sums1[0].final_cost:.9e} cams rel {rel(cams, c1):.2e} pts rel {rel(pts, p1):.2e}"
assert rc == 0 and rc1 == 0 and acc == acc1, (msg, acc, acc1)
assert rel(cams, c1) < 1e-6 and rel(pts, p1) < 1e-6, msg
if rank == 0:
    import oracle_binding as ob
    ob.lib().uba_ref_set_threads(os.cpu_count() or 1)
    o = ob.optimise(win, capi.default_config(loss_kind=loss, fixed_iterations=K), 2)
    acco = [b["accepted"] for b in o["iterations"]]
    msg += f" | oracle: cams rel {rel(cams, o['cams']):.2e} pts rel {rel(pts, o['pts'][ids]):.2e}"
    assert acc == acco and rel(cams, o["cams"]) < 1e-6 and rel(pts, o["pts"][ids]) < 1e-6, msg
print(msg, flush=True)
h.close(); h1.close()
dist.barrier()

# (2) the reference's termination rules, no fixed K (every rank must leave the loop in the same iteration)
h = sharded_handle(max_solver_time_s=0.0)
rc, sums = h.optimise(2)
h1 = single_handle(max_solver_time_s=0.0)
rc1, sums1 = h1.optimise(2)
t = torch.tensor([sums[0].iterations, sums[0].termination], dtype=torch.int64, device="cuda")
tl = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(tl, t)
assert all(bool((x == t).all()) for x in tl), [x.tolist() for x in tl]
assert rc == 0 and sums[0].iterations == sums1[0].iterations and sums[0].termination == sums1[0].termination, (sums[0].iterations, sums1[0].iterations)
assert rel(h.cameras(), h1.cameras()) < 1e-6
print(f"rank {rank}/{world}: free-running: {sums[0].iterations} iterations, termination {sums[0].termination} on every rank", flush=True)
h.close(); h1.close()
dist.barrier()

# (3) a 1 ms wall-clock cap: ranks have their own clocks, the stop must still be taken together
h = sharded_handle(max_solver_time_s=1e-3, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0, max_iterations=30)
rc, sums = h.optimise(2)
t = torch.tensor([sums[0].iterations], dtype=torch.int64, device="cuda")
tl = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(tl, t)
assert rc == 0 and all(int(x) == int(t) for x in tl) and int(t) < 30, [int(x) for x in tl]
print(f"rank {rank}/{world}: wall-clock cap: stopped together after {int(t)} iterations", flush=True)
h.close()
dist.barrier()

# (4) pose covariances on a sharded handle
if scale <= 0.25:
    h = sharded_handle(fixed_iterations=3, compute_covariance=1)
    h.optimise(2)
    h1 = single_handle(fixed_iterations=3, compute_covariance=1)
    h1.optimise(2)
    cv, cv1 = h.pose_covariances(), h1.pose_covariances()
    assert rel(cv, cv1) < 1e-7, rel(cv, cv1)
    print(f"rank {rank}/{world}: pose covariances rel {rel(cv, cv1):.2e}", flush=True)
    h.close(); h1.close()
dist.barrier()
if rank == 0:
    print("MULTIGPU CHECK OK", flush=True)
dist.destroy_process_group()
