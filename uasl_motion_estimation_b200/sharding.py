"""Point sharding of one large window across ranks — ctypes plumbing over the C ABI.

The partition itself lives behind ``uba_shard_points`` / ``uba_shard_extract`` (include/uba.h; host-only code, carried by
libuba_host.so and libuba.so) so that the C++ drop-in header can shard the same way.  Points are assigned to ranks by
KEYFRAME RANGE: ordered by (first keyframe of the track, caller index) and cut into pieces of equal observation count.
Every rank keeps ALL cameras; a rank's Schur products touch one stretch of the block band of the reduced camera system,
which is the only thing that crosses ranks (summed over NVLink peer memory inside libuba).  Batches of independent
windows shard by window, no collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .synth import Window


def point_ranks(win: Window, n_ranks: int, lib=None):
    """(pt_rank [n_pts], rank_obs [n_ranks], rank_pts [n_ranks])."""
    lib = lib or capi.host_lib()
    pt_rank = np.zeros(win.n_pts, np.int32); rank_obs = np.zeros(n_ranks, np.int64); rank_pts = np.zeros(n_ranks, np.int32)
    rc = lib.uba_shard_points(win.n_cams, win.n_pts, win.n_obs, capi.i32ptr(win.cam_idx), capi.i32ptr(win.pt_idx), n_ranks,
                              capi.i32ptr(pt_rank), capi.i64ptr(rank_obs), capi.i32ptr(rank_pts))
    if rc != 0:
        raise capi.UbaError(rc, "uba_shard_points failed")
    return pt_rank, rank_obs, rank_pts


def shard_window(win: Window, rank: int, n_ranks: int, lib=None, return_ids: bool = False):
    """This rank's shard: all cameras, the points whose tracks start in the rank's keyframe range, their observations."""
    lib = lib or capi.host_lib()
    pt_rank, rank_obs, rank_pts = point_ranks(win, n_ranks, lib)
    npt, no, M = int(rank_pts[rank]), int(rank_obs[rank]), win.M
    pts = np.zeros((npt, 3)); feats = np.zeros((no, M)); ci = np.zeros(no, np.int32); pi = np.zeros(no, np.int32)
    cid = np.zeros(no, np.int32); ids = np.zeros(npt, np.int32)
    n = lib.uba_shard_extract(M, win.n_pts, win.n_obs, capi.dptr(win.pts_init), capi.dptr(win.feats), capi.i32ptr(win.cam_idx),
                              capi.i32ptr(win.pt_idx), capi.i32ptr(win.cam_id), capi.i32ptr(pt_rank), rank, capi.dptr(pts),
                              capi.dptr(feats), capi.i32ptr(ci), capi.i32ptr(pi), capi.i32ptr(cid), capi.i32ptr(ids))
    if n != no:
        raise capi.UbaError(int(n), "uba_shard_extract failed")
    shard = Window(M, win.cams_gt, win.cams_init, np.ascontiguousarray(win.pts_gt[ids]), pts, feats, ci, pi, cid, win.fixed_frames, win.calib)
    return (shard, ids) if return_ids else shard


def window_ranges(n_windows: int, n_ranks: int) -> np.ndarray:
    """[n_ranks+1] window boundaries for a batch of independent windows."""
    return np.array([(n_windows * r) // n_ranks for r in range(n_ranks + 1)], np.int64)
