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
    lib = capi.load_host()  # generator, defaults and shard tables: the host-only library
    win = synth.config_window("c4", scale=0.003, lib=lib)
    cfg = capi.default_config(lib)
    full = ob.linearize(win, cfg, 2, 1e4)
    sh = sharding.shard_window(win, rank, world, lib)
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
        out["balance"] = [int(x) for x in sharding.point_ranks(win, world, lib)[1]]
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


def test_keyframe_range_shards_cover_balance_and_stay_local():
    """uba_shard_points: every point lands on exactly one rank, observation counts are balanced, a rank's tracks START
    inside one contiguous keyframe range (so only neighbouring ranks share cameras), and the extracted shards reassemble
    the window."""
    win = synth.config_window("c4", scale=0.02)
    for world in (1, 2, 3, 8):
        pt_rank, rank_obs, rank_pts = sharding.point_ranks(win, world)
        assert pt_rank.min() >= 0 and pt_rank.max() < world
        assert rank_obs.sum() == win.n_obs and rank_pts.sum() == win.n_pts
        assert rank_obs.max() - rank_obs.min() <= 8 * world
        lo = np.full(win.n_pts, win.n_cams); np.minimum.at(lo, win.pt_idx, win.cam_idx)
        first = [lo[(pt_rank == r) & (lo < win.n_cams)] for r in range(world)]
        for r in range(world - 1):
            if len(first[r]) and len(first[r + 1]):
                assert first[r].max() <= first[r + 1].min()        # keyframe ranges are ordered, overlapping in one keyframe at most
        seen_pts = np.zeros(win.n_pts, int); n_obs = 0
        for r in range(world):
            sh, ids = sharding.shard_window(win, r, world, return_ids=True)
            seen_pts[ids] += 1; n_obs += sh.n_obs
            assert np.array_equal(sh.pts_init, win.pts_init[ids])
            sel = pt_rank[win.pt_idx] == r
            assert np.array_equal(sh.feats, win.feats[sel]) and np.array_equal(sh.cam_idx, win.cam_idx[sel])
            assert np.array_equal(ids[sh.pt_idx], win.pt_idx[sel])
        assert (seen_pts == 1).all() and n_obs == win.n_obs


def test_window_ranges():
    assert list(sharding.window_ranges(4096, 8)) == [512 * r for r in range(9)]
    assert list(sharding.window_ranges(10, 4)) == [0, 2, 5, 7, 10]
