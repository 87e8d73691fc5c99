"""Point sharding on the CPU: world size 2 over gloo.  Every rank linearises ITS
point shard with the oracle, the reduced camera systems are summed with torch.distributed, and the sum
must equal the single-process system — the identity libuba's NCCL allreduce relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uasl_motion_estimation_b200 import capi, sharding, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, emu_path, out):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import oracle_binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = capi.load(emu_path)  # only for the generator + defaults (host code)
    win = synth.config_window("c4", scale=0.003, lib=lib)
    cfg = capi.default_config(lib)
    full = ob.linearize(win, cfg, 2, 1e4)
    sh = sharding.shard_window(win, rank, world)
    # Jacobi scaling and camera damping are global quantities: linearise undamped and add them after the reduce
    part = ob.linearize(sh, cfg, 2, -1.0, jacobi_scale=np.ones(6 * sh.n_cams + 3 * sh.n_pts))
    nfree = int((ob.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, 2)["free_cam"] >= 0).sum())
    local_free = ob.tables(sh.n_cams, sh.n_pts, sh.cam_idx, sh.pt_idx, 2)["free_cam"]
    # scatter the shard's (smaller) reduced system into the global camera numbering
    S = np.zeros((6 * nfree, 6 * nfree)); B = torch.from_numpy(part["B"].copy()); g = torch.from_numpy(part["grad_cams"].copy())
    gfree = ob.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, 2)["free_cam"]
    idx = [c for c in range(win.n_cams) if local_free[c] >= 0]
    rows = np.concatenate([np.arange(6 * gfree[c], 6 * gfree[c] + 6) for c in idx]) if idx else np.zeros(0, int)
    S[np.ix_(rows, rows)] = part["S"]
    cost = torch.tensor([part["cost"][0]])
    St = torch.from_numpy(S)
    for tns in (St, B, g, cost):
        dist.all_reduce(tns)
    if rank == 0:
        out["cost"] = (cost.item(), full["cost"][0])
        out["B"] = np.abs(B.numpy() - full["B"]).max() / np.abs(full["B"]).max()
        out["g"] = np.abs(g.numpy() - full["grad_cams"]).max() / np.abs(full["grad_cams"]).max()
        # undamped point blocks differ from the damped full run only through lambda: compare against an undamped full run
        full_u = ob.linearize(win, cfg, 2, -1.0, jacobi_scale=np.ones(6 * win.n_cams + 3 * win.n_pts))
        out["S"] = np.abs(St.numpy() - full_u["S"]).max() / np.abs(full_u["S"]).max()
        b = sharding.point_ranges(win.pt_idx, win.n_pts, world)
        out["balance"] = [int(((win.pt_idx >= b[r]) & (win.pt_idx < b[r + 1])).sum()) for r in range(world)]
    dist.destroy_process_group()


def test_point_sharded_reduced_system_sums_to_the_full_one(emu_lib):
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    emu_path = str(capi.PKG_DIR.parent / "tests" / "emu" / "libuba_emu.so")
    mp.spawn(_worker, args=(world, _free_port(), emu_path, out), nprocs=world, join=True)
    assert out["cost"][0] == pytest.approx(out["cost"][1], rel=1e-12)
    assert out["B"] < 1e-12 and out["g"] < 1e-12 and out["S"] < 1e-10
    n = out["balance"]
    assert abs(n[0] - n[1]) <= 0.02 * sum(n)  # balanced by observation count


def test_point_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    k = rng.integers(0, 12, size=5000)
    pt_idx = np.repeat(np.arange(5000), k)
    for world in (1, 2, 3, 8):
        b = sharding.point_ranges(pt_idx, 5000, world)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all()
        counts = [int(((pt_idx >= b[r]) & (pt_idx < b[r + 1])).sum()) for r in range(world)]
        assert sum(counts) == len(pt_idx) and max(counts) - min(counts) <= 12 * world


def test_window_ranges():
    assert list(sharding.window_ranges(4096, 8)) == [512 * r for r in range(9)]
    assert list(sharding.window_ranges(10, 4)) == [0, 2, 5, 7, 10]
