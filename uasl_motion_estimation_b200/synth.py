"""Synthetic stereo-rig windows of BASELINE.json's five configurations.

Thin wrapper over the C generator ``uba_synth_generate`` (csrc/uba_synth.cpp): seeds are
``20261018 + 1000*config + window`` and every array comes back in the caller order the
reference's ``initialiseObservations`` produces (BundleAdjuster.h:364-374): point-major,
frame-ascending.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi

BASE_SEED = 20261018


@dataclass
class Window:
    M: int
    cams_gt: np.ndarray
    cams_init: np.ndarray
    pts_gt: np.ndarray
    pts_init: np.ndarray
    feats: np.ndarray
    cam_idx: np.ndarray
    pt_idx: np.ndarray
    cam_id: np.ndarray
    fixed_frames: int
    calib: capi.Calib

    @property
    def n_cams(self):
        return self.cams_init.shape[0]

    @property
    def n_pts(self):
        return self.pts_init.shape[0]

    @property
    def n_obs(self):
        return self.feats.shape[0]


# name -> (config id, n_cams, n_pts, track_min, track_max, full_tracks, outlier_fraction, loss, fixed LM iterations)
CONFIGS = {
    "c1": dict(config=1, n_cams=10, n_pts=2000, track_min=10, track_max=10, full_tracks=1, outliers=0.0, loss=capi.LOSS_HUBER, iters=10),
    "c2": dict(config=2, n_cams=20, n_pts=20000, track_min=2, track_max=18, full_tracks=0, outliers=0.0, loss=capi.LOSS_HUBER, iters=10),
    "c3": dict(config=3, n_cams=10, n_pts=2000, track_min=10, track_max=10, full_tracks=1, outliers=0.0, loss=capi.LOSS_HUBER, iters=10, windows=4096),
    "c4": dict(config=4, n_cams=200, n_pts=200000, track_min=5, track_max=5, full_tracks=0, outliers=0.0, loss=capi.LOSS_HUBER, iters=10),
    "c5": dict(config=5, n_cams=100, n_pts=100000, track_min=5, track_max=5, full_tracks=0, outliers=0.3, loss=capi.LOSS_CAUCHY, iters=20),
}


def generate(n_cams, n_pts, track_min, track_max, full_tracks=0, outliers=0.0, seed=BASE_SEED, M=4, fixed_frames=2,
             calib=None, lib=None, pixel_sigma=0.5, pose_t_sigma=0.05, pose_r_sigma=0.005, point_rel_sigma=0.01) -> Window:
    lib = lib or capi.host_lib()
    calib = calib or capi.default_calib(lib)
    spec = capi.SynthSpec(M=M, n_cams=n_cams, n_pts=n_pts, track_min=track_min, track_max=track_max, full_tracks=full_tracks,
                          outlier_fraction=outliers, pixel_sigma=pixel_sigma, pose_t_sigma=pose_t_sigma,
                          pose_r_sigma=pose_r_sigma, point_rel_sigma=point_rel_sigma, fixed_frames=fixed_frames, seed=seed)
    max_obs = int(n_pts) * int(n_cams if full_tracks else min(track_max, n_cams))
    cams_gt = np.zeros((n_cams, 6)); cams_init = np.zeros((n_cams, 6))
    pts_gt = np.zeros((n_pts, 3)); pts_init = np.zeros((n_pts, 3))
    feats = np.zeros((max(max_obs, 1), M)); cam_idx = np.zeros(max(max_obs, 1), np.int32)
    pt_idx = np.zeros(max(max_obs, 1), np.int32); cam_id = np.zeros(max(max_obs, 1), np.int32)
    n = lib.uba_synth_generate(C.byref(spec), C.byref(calib), max_obs, capi.dptr(cams_gt), capi.dptr(cams_init), capi.dptr(pts_gt),
                               capi.dptr(pts_init), capi.dptr(feats), capi.i32ptr(cam_idx), capi.i32ptr(pt_idx), capi.i32ptr(cam_id))
    if n < 0:
        raise capi.UbaError(int(n), "uba_synth_generate failed")
    n = int(n)
    return Window(M, cams_gt, cams_init, pts_gt, pts_init, np.ascontiguousarray(feats[:n]), np.ascontiguousarray(cam_idx[:n]),
                  np.ascontiguousarray(pt_idx[:n]), np.ascontiguousarray(cam_id[:n]), fixed_frames, calib)


def reorder(win: Window, point_order: np.ndarray | None = None, camera_descending: bool = False) -> Window:
    """The same window with its points renumbered (`point_order[new] = old`) and its observations re-emitted point-major —
    camera-ascending inside a point as the reference's initialiseObservations does (BundleAdjuster.h:364-374), or descending.
    `point_order=None` sorts the tracks by (first, last) keyframe: the order a sliding window's track container has."""
    counts = np.bincount(win.pt_idx, minlength=win.n_pts)
    lo = np.full(win.n_pts, np.iinfo(np.int32).max, np.int64); hi = np.full(win.n_pts, -1, np.int64)
    np.minimum.at(lo, win.pt_idx, win.cam_idx); np.maximum.at(hi, win.pt_idx, win.cam_idx)
    if point_order is None:
        point_order = np.lexsort((hi, np.where(counts > 0, lo, np.iinfo(np.int32).max)))
    new_of_old = np.empty(win.n_pts, np.int64); new_of_old[point_order] = np.arange(win.n_pts)
    new_pt = new_of_old[win.pt_idx]
    key_cam = -win.cam_idx.astype(np.int64) if camera_descending else win.cam_idx.astype(np.int64)
    o = np.lexsort((key_cam, new_pt))
    return Window(win.M, win.cams_gt, win.cams_init, np.ascontiguousarray(win.pts_gt[point_order]), np.ascontiguousarray(win.pts_init[point_order]),
                  np.ascontiguousarray(win.feats[o]), np.ascontiguousarray(win.cam_idx[o]), np.ascontiguousarray(new_pt[o].astype(np.int32)),
                  np.ascontiguousarray(win.cam_id[o]), win.fixed_frames, win.calib)


def config_window(name: str, window: int = 0, scale: float = 1.0, lib=None, M=4) -> Window:
    """One window of configuration c1..c5; ``scale`` shrinks the point count (parity tests)."""
    c = CONFIGS[name]
    n_pts = max(8, int(round(c["n_pts"] * scale)))
    return generate(c["n_cams"], n_pts, c["track_min"], c["track_max"], c["full_tracks"], c["outliers"],
                    seed=BASE_SEED + 1000 * c["config"] + window, M=M, lib=lib)


def concat_windows(wins) -> dict:
    """Concatenate windows into uba_set_batch arguments."""
    wc = np.zeros(len(wins) + 1, np.int32); wp = np.zeros(len(wins) + 1, np.int32); wo = np.zeros(len(wins) + 1, np.int64)
    for i, w in enumerate(wins):
        wc[i + 1] = wc[i] + w.n_cams; wp[i + 1] = wp[i] + w.n_pts; wo[i + 1] = wo[i] + w.n_obs
    return dict(M=wins[0].M, win_cam_off=wc, win_pt_off=wp, win_obs_off=wo,
                cams6=np.concatenate([w.cams_init for w in wins]), pts3=np.concatenate([w.pts_init for w in wins]),
                feats=np.concatenate([w.feats for w in wins]), cam_idx=np.concatenate([w.cam_idx for w in wins]),
                pt_idx=np.concatenate([w.pt_idx for w in wins]), cam_id=np.concatenate([w.cam_id for w in wins]),
                calib=wins[0].calib)


def vo_params(calib, lib=None, **overrides) -> "capi.VoParams":
    """StereoVisualOdometry::parameters for the synthetic rig (defaults of the reference, vo/VisualOdometry.h:32)."""
    lib = lib or capi.host_lib()
    p = capi.VoParams()
    lib.uba_vo_params_default(C.byref(p))
    p.fu1, p.fv1, p.cu1, p.cv1 = calib.fx0, calib.fy0, calib.cx0, calib.cy0
    p.fu2, p.fv2, p.cu2, p.cv2 = calib.fx1, calib.fy1, calib.cx1, calib.cy1
    p.baseline = calib.baseline
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def vo_quads(n_matches: int, outlier_fraction: float = 0.0, seed: int = BASE_SEED + 7000, lib=None):
    """Quad matches of two consecutive stereo frames (StereoOdoMatchesf f1..f4): the two keyframes of a synthetic 2-frame
    window, points seen in both; a fraction of the CURRENT-frame features is replaced by uniform-random pixels.
    Returns (quads [n][8] float32, outlier mask)."""
    win = generate(2, n_matches, 2, 2, full_tracks=1, seed=seed, lib=lib)
    both = np.flatnonzero(np.bincount(win.pt_idx, minlength=win.n_pts) == 2)     # tracks that stayed inside both images
    first = np.searchsorted(win.pt_idx, both)                                   # observations are point-major, frame-ascending
    quads = np.ascontiguousarray(np.concatenate([win.feats[first], win.feats[first + 1]], axis=1), dtype=np.float32)
    rng = np.random.default_rng(seed)
    out = rng.random(len(quads)) < outlier_fraction
    k = int(out.sum())
    quads[out, 4] = rng.uniform(20, 1221, k); quads[out, 5] = rng.uniform(20, 356, k)
    quads[out, 6] = quads[out, 4] - rng.uniform(2, 90, k).astype(np.float32); quads[out, 7] = quads[out, 5]
    return quads, out


class SlidingSequence:
    """A long synthetic stereo sequence walked by a sliding window — BASELINE config c2: 200 keyframes,
    a 20-keyframe window advanced one keyframe per BA call (181 calls).  It plays the caller's part around
    ``uba_window_advance`` the way an application drives the reference (``BundleAdjuster.h:351-376`` over a container of
    ``WBA_Point`` tracks, ``core/feature_types.h:121-191``): a track enters the container at its first keyframe, gains one
    observation per keyframe while it is matched, and is erased once its last observation has left the window.

    Tracks are cut at their first gap (a tracker loses the feature there), so every track is a run of consecutive
    keyframes, which is what ``WBA_Point::addMatch`` asserts."""

    def __init__(self, n_frames=200, window=20, n_pts=200_000, track_min=2, track_max=18, seed=BASE_SEED + 2000, lib=None):
        self.W = window
        self.n_frames = n_frames
        seq = generate(n_frames, n_pts, track_min, track_max, 0, 0.0, seed=seed, lib=lib)
        # keep the leading run of consecutive keyframes of every track
        first = np.r_[True, seq.pt_idx[1:] != seq.pt_idx[:-1]]
        brk = ~first & (seq.cam_idx != np.r_[0, seq.cam_idx[:-1]] + 1)
        seg = np.cumsum(first) - 1
        nbrk = np.cumsum(brk)
        keep = nbrk == nbrk[np.flatnonzero(first)][seg]      # no gap yet inside this track
        self.seq = seq
        self.cam = seq.cam_idx[keep].astype(np.int64); self.pt = seq.pt_idx[keep].astype(np.int64)
        self.feats = np.ascontiguousarray(seq.feats[keep])
        self.lo = np.full(seq.n_pts, np.iinfo(np.int64).max); self.hi = np.full(seq.n_pts, -1)
        np.minimum.at(self.lo, self.pt, self.cam); np.maximum.at(self.hi, self.pt, self.cam)
        self.off = np.zeros(seq.n_pts + 1, np.int64); np.cumsum(np.bincount(self.pt, minlength=seq.n_pts), out=self.off[1:])
        self.calib = seq.calib

    @property
    def n_calls(self):
        return self.n_frames - self.W + 1

    def initial_ids(self):
        """Tracks alive in the first window, in container order."""
        return np.flatnonzero((self.hi >= 0) & (self.lo < self.W))

    def window(self, first, ids, cams=None, pts=None) -> Window:
        """The window [first, first + W) over the tracks `ids` (container order) as the arrays initialiseObservations
        would produce from them; initial values from `cams` / `pts` or the generator's perturbed ones."""
        f1 = first + self.W
        a = np.maximum(self.lo[ids], first); b = np.minimum(self.hi[ids], f1 - 1)
        cnt = np.maximum(b - a + 1, 0)
        n = int(cnt.sum())
        pt_new = np.repeat(np.arange(len(ids)), cnt)
        start = np.zeros(len(ids) + 1, np.int64); np.cumsum(cnt, out=start[1:])
        q = np.arange(n) - start[pt_new]
        src = self.off[ids][pt_new] + (a - self.lo[ids])[pt_new] + q
        return Window(4, self.seq.cams_gt[first:f1], self.seq.cams_init[first:f1] if cams is None else cams,
                      self.seq.pts_gt[ids], self.seq.pts_init[ids] if pts is None else pts, np.ascontiguousarray(self.feats[src]),
                      (self.cam[src] - first).astype(np.int32), pt_new.astype(np.int32), np.zeros(n, np.int32), 2, self.calib)

    def advance(self, first, ids):
        """One step of the window, [first, first + W) -> [first + 1, first + 1 + W): the arguments of uba_window_advance and
        the container afterwards.  Returns (kwargs, ids_new)."""
        newf = first + self.W                      # the keyframe that enters
        alive = self.hi[ids] >= first + 1
        born = np.flatnonzero(self.lo == newf)
        ids_new = np.concatenate([ids[alive], born])
        seen = np.flatnonzero((self.lo[ids_new] <= newf) & (self.hi[ids_new] >= newf))   # tracks observed in the new keyframe
        src = self.off[ids_new[seen]] + (newf - self.lo[ids_new[seen]])
        kw = dict(n_drop=1, new_cams6=self.seq.cams_init[newf:newf + 1], new_pts3=self.seq.pts_init[born],
                  feats=np.ascontiguousarray(self.feats[src]), cam_idx=np.full(len(seen), self.W - 1, np.int32), pt_idx=seen.astype(np.int32))
        return kw, ids_new
