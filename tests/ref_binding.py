"""ctypes binding of oracle/_ref/libuba_ref.so: the REFERENCE'S OWN SOURCES (BundleAdjuster.h, rotation_utils.cpp,
StereoVisualOdometry.cpp) compiled against the minimal Ceres / OpenCV stand-ins of oracle/refstub.  TEST INFRASTRUCTURE.
The library is built by `make -C oracle ref` where /root/reference exists and travels to the GPU box prebuilt."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from uasl_motion_estimation_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
REF_LIB = ROOT / "oracle" / "_ref" / "libuba_ref.so"
dp, ip, lp = capi.c_double_p, capi.c_int32_p, capi.c_int64_p
_lib = None


def available() -> bool:
    if not REF_LIB.exists() and Path("/root/reference/include/MotionEstimation").is_dir():
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "ref"], check=False, capture_output=True)
    return REF_LIB.exists()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError(f"{REF_LIB} not built (needs /root/reference): make -C oracle ref")
        L = C.CDLL(str(REF_LIB))
        L.uba_refsrc_residual.restype = C.c_int
        L.uba_refsrc_residual.argtypes = [C.c_int, C.POINTER(capi.Calib), dp, dp, dp, C.c_int, dp, dp, dp]
        L.uba_refsrc_log_map_quat.argtypes = [dp, dp]; L.uba_refsrc_exp_map_quat.argtypes = [dp, dp]
        L.uba_refsrc_ba_run.restype = C.c_int
        L.uba_refsrc_ba_run.argtypes = [C.c_int, C.POINTER(capi.Calib), C.c_int, C.c_int, dp, ip, C.c_int, dp, ip, lp, ip, dp, C.c_int, C.c_int,
                                        C.c_int64, lp, ip, ip, ip, dp, dp, dp, dp, dp, ip, dp]
        L.uba_refsrc_vo_project3d.restype = C.c_int
        L.uba_refsrc_vo_project3d.argtypes = [dp, C.c_int, dp, dp]
        L.uba_refsrc_vo_linearize.restype = C.c_int
        L.uba_refsrc_vo_linearize.argtypes = [dp, C.c_int, dp, dp, C.c_int, ip, dp, dp, dp, dp, dp]
        L.uba_refsrc_vo_inliers.restype = C.c_int
        L.uba_refsrc_vo_inliers.argtypes = [dp, C.c_int, dp, dp, ip]
        L.uba_refsrc_vo_optimize.restype = C.c_int
        L.uba_refsrc_vo_optimize.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, ip, dp, ip, ip]
        L.uba_refsrc_vo_process.restype = C.c_int
        L.uba_refsrc_vo_process.argtypes = [dp, dp, C.c_int, C.c_uint, C.c_int, dp, dp, ip, ip]
        _lib = L
    return _lib


def residual(M, calib, cam6, pt3, obs, cam_id=0, jac=True):
    cam6 = capi.as_f64(cam6); pt3 = capi.as_f64(pt3); obs = capi.as_f64(obs)
    r = np.zeros(M); Jc = np.zeros((M, 6)); Jp = np.zeros((M, 3))
    rc = lib().uba_refsrc_residual(M, C.byref(calib), capi.dptr(cam6), capi.dptr(pt3), capi.dptr(obs), cam_id, capi.dptr(r),
                                   capi.dptr(Jc) if jac else None, capi.dptr(Jp) if jac else None)
    assert rc == 0
    return r, Jc, Jp


def log_map(q):
    q = capi.as_f64(q); r = np.zeros(3); lib().uba_refsrc_log_map_quat(capi.dptr(q), capi.dptr(r)); return r


def exp_map(r):
    r = capi.as_f64(r); q = np.zeros(4); lib().uba_refsrc_exp_map_quat(capi.dptr(r), capi.dptr(q)); return q


def ba_run(win, first_frame=100, fixed_frames=2, optimise=True, compute_cov=False, pts_from_tracks=True):
    """Feeds a synthetic window to the reference's BundleAdjuster<M> the way its callers do: CamPose_qd poses with frame IDs
    first_frame.., one WBA point per track carrying its 3-D location (homogeneous) — the points are initialised from the
    tracks (BundleAdjuster.h:365-367)."""
    M, nc, npt = win.M, win.n_cams, win.n_pts
    poses = np.zeros((nc, 7))
    for c in range(nc):
        poses[c, :4] = exp_map(win.cams_init[c, 3:]); poses[c, 4:] = win.cams_init[c, :3]
    cam_ids = np.arange(first_frame, first_frame + nc, dtype=np.int32)
    counts = np.bincount(win.pt_idx, minlength=npt)
    track_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    order = np.argsort(win.pt_idx, kind="stable")
    frame_idx = (win.cam_idx[order] + first_frame).astype(np.int32)
    feats = np.ascontiguousarray(win.feats[order])
    pt_cam = np.zeros(npt, np.int32)
    first = np.full(npt, -1, np.int64); first[win.pt_idx[order][::-1]] = np.arange(len(order))[::-1]
    pt_cam[counts > 0] = win.cam_id[order][first[counts > 0]]
    pts4 = np.concatenate([win.pts_init, np.ones((npt, 1))], axis=1)
    no = C.c_int64(0); max_obs = win.n_obs + 8
    out = dict(cam_idx=np.zeros(max_obs, np.int32), pt_idx=np.zeros(max_obs, np.int32), cam_id=np.zeros(max_obs, np.int32),
               feats=np.zeros((max_obs, M)), cams_init=np.zeros((nc, 6)), cams=np.zeros((nc, 6)), pts=np.zeros((npt, 3)),
               quat_id=np.zeros((nc, 5)), cov=np.zeros((nc, 6, 6)))
    status = C.c_int32(-1)
    rc = lib().uba_refsrc_ba_run(M, C.byref(win.calib), int(compute_cov), nc, capi.dptr(poses), capi.i32ptr(cam_ids), npt, capi.dptr(pts4),
                                 capi.i32ptr(pt_cam), capi.i64ptr(track_off), capi.i32ptr(frame_idx), capi.dptr(feats), fixed_frames,
                                 int(optimise), max_obs, C.byref(no), capi.i32ptr(out["cam_idx"]), capi.i32ptr(out["pt_idx"]),
                                 capi.i32ptr(out["cam_id"]), capi.dptr(out["feats"]), capi.dptr(out["cams_init"]), capi.dptr(out["cams"]),
                                 capi.dptr(out["pts"]), capi.dptr(out["quat_id"]), C.byref(status), capi.dptr(out["cov"]))
    n = no.value
    for k in ("cam_idx", "pt_idx", "cam_id", "feats"):
        out[k] = out[k][:n]
    out.update(rc=rc, status=status.value, n_obs=n)
    return out


# ---- stereo VO ----
def vo_params(calib, inlier_threshold=2.0):
    return np.array([calib.fx0, calib.fy0, calib.cx0, calib.cy0, calib.fx1, calib.fy1, calib.cx1, calib.cy1, calib.baseline, inlier_threshold])


def vo_project3d(p10, quads):
    quads = capi.as_f64(quads); out = np.zeros((len(quads), 4))
    lib().uba_refsrc_vo_project3d(capi.dptr(p10), len(quads), capi.dptr(quads), capi.dptr(out))
    return out


def vo_linearize(p10, quads, state, selection):
    quads = capi.as_f64(quads); state = capi.as_f64(state); sel = capi.as_i32(selection); n = len(sel)
    pred = np.zeros((n, 4)); res = np.zeros((n, 4)); J = np.zeros((6, 4 * n)); A = np.zeros((6, 6)); B = np.zeros(6)
    lib().uba_refsrc_vo_linearize(capi.dptr(p10), len(quads), capi.dptr(quads), capi.dptr(state), n, capi.i32ptr(sel), capi.dptr(pred),
                                  capi.dptr(res), capi.dptr(J), capi.dptr(A), capi.dptr(B))
    return dict(pred=pred, res=res, J=J, A=A, B=B)


def vo_inliers(p10, quads, state):
    quads = capi.as_f64(quads); state = capi.as_f64(state); inl = np.zeros(len(quads), np.int32)
    n = lib().uba_refsrc_vo_inliers(capi.dptr(p10), len(quads), capi.dptr(quads), capi.dptr(state), capi.i32ptr(inl))
    return inl[:n].copy()


def vo_optimize(p10, opt6, quads, state, selection):
    quads = capi.as_f64(quads); state = capi.as_f64(state); sel = capi.as_i32(selection); opt6 = capi.as_f64(opt6)
    out = np.zeros(6); inl = np.zeros(len(quads), np.int32); n_in = C.c_int32(0)
    ok = lib().uba_refsrc_vo_optimize(capi.dptr(p10), capi.dptr(opt6), len(quads), capi.dptr(quads), capi.dptr(state), len(sel),
                                      capi.i32ptr(sel), capi.dptr(out), capi.i32ptr(inl), C.byref(n_in))
    return bool(ok), out, inl[:n_in.value].copy()


def vo_process(p10, opt6, n_ransac, seed, quads):
    quads = capi.as_f64(quads); opt6 = capi.as_f64(opt6)
    out = np.zeros(6); inl = np.zeros(len(quads), np.int32); n_in = C.c_int32(0)
    ok = lib().uba_refsrc_vo_process(capi.dptr(p10), capi.dptr(opt6), n_ransac, seed, len(quads), capi.dptr(quads), capi.dptr(out),
                                     capi.i32ptr(inl), C.byref(n_in))
    return bool(ok), out, inl[:n_in.value].copy()
