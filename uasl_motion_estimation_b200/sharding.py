"""Point sharding of one large window across ranks.

Observations are point-major (BundleAdjuster.h:364-374), so a contiguous point range is a
contiguous observation range.  Every rank keeps ALL cameras and a contiguous range of points
balanced by observation count; the reduced camera system is the only thing that crosses ranks
(NCCL allreduce inside libuba).  Batches of independent windows shard by window, no collective.
"""
from __future__ import annotations

import numpy as np

from .synth import Window


def point_ranges(pt_idx: np.ndarray, n_pts: int, n_ranks: int) -> np.ndarray:
    """[n_ranks+1] point boundaries: contiguous ranges with (nearly) equal observation counts."""
    counts = np.bincount(np.asarray(pt_idx, dtype=np.int64), minlength=n_pts)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    bounds = np.zeros(n_ranks + 1, np.int64)
    for r in range(1, n_ranks):
        bounds[r] = np.searchsorted(csum, total * r / n_ranks, side="left")
    bounds[n_ranks] = n_pts
    return np.maximum.accumulate(bounds)


def shard_window(win: Window, rank: int, n_ranks: int) -> Window:
    """This rank's shard: all cameras, points [b[rank], b[rank+1]) and their observations."""
    b = point_ranges(win.pt_idx, win.n_pts, n_ranks)
    lo, hi = int(b[rank]), int(b[rank + 1])
    sel = (win.pt_idx >= lo) & (win.pt_idx < hi)
    return Window(win.M, win.cams_gt, win.cams_init, win.pts_gt[lo:hi], np.ascontiguousarray(win.pts_init[lo:hi]),
                  np.ascontiguousarray(win.feats[sel]), np.ascontiguousarray(win.cam_idx[sel]),
                  np.ascontiguousarray(win.pt_idx[sel] - lo).astype(np.int32), np.ascontiguousarray(win.cam_id[sel]),
                  win.fixed_frames, win.calib)


def window_ranges(n_windows: int, n_ranks: int) -> np.ndarray:
    """[n_ranks+1] window boundaries for a batch of independent windows."""
    return np.array([(n_windows * r) // n_ranks for r in range(n_ranks + 1)], np.int64)
