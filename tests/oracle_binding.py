"""ctypes binding of the CPU oracle (oracle/libuba_oracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from uasl_motion_estimation_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "libuba_oracle.so"

dp, ip, lp = capi.c_double_p, capi.c_int32_p, capi.c_int64_p


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not ORACLE_LIB.exists():
            build()
        L = C.CDLL(str(ORACLE_LIB))
        L.uba_ref_residual.restype = C.c_int
        L.uba_ref_residual.argtypes = [C.c_int, C.POINTER(capi.Calib), dp, dp, dp, C.c_int, dp, dp, dp]
        L.uba_ref_rotate.restype = None
        L.uba_ref_rotate.argtypes = [dp, dp, dp]
        L.uba_ref_loss.restype = None
        L.uba_ref_loss.argtypes = [C.c_int, C.c_double, C.c_double, dp]
        L.uba_ref_point_bounds.restype = None
        L.uba_ref_point_bounds.argtypes = [C.POINTER(capi.Calib), C.c_int, dp, dp]
        L.uba_ref_log_map_quat.restype = None
        L.uba_ref_log_map_quat.argtypes = [dp, dp]
        L.uba_ref_exp_map_quat.restype = None
        L.uba_ref_exp_map_quat.argtypes = [dp, dp]
        L.uba_ref_tables.restype = C.c_int
        L.uba_ref_tables.argtypes = [C.c_int, C.c_int, C.c_int64, ip, ip, C.c_int, ip, lp, ip, ip]
        L.uba_ref_linearize.restype = C.c_int
        L.uba_ref_linearize.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, dp, ip, ip, ip, C.POINTER(capi.Calib),
                                        C.POINTER(capi.Config), C.c_int, C.c_double, dp, C.POINTER(capi.LinearizationOut), dp, dp, dp, dp]
        L.uba_ref_cost.restype = C.c_int
        L.uba_ref_cost.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, dp, ip, ip, ip, C.POINTER(capi.Calib),
                                   C.POINTER(capi.Config), dp]
        L.uba_ref_optimise.restype = C.c_int
        L.uba_ref_optimise.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, dp, ip, ip, ip, C.POINTER(capi.Calib),
                                       C.POINTER(capi.Config), C.c_int, C.POINTER(capi.Summary), C.POINTER(capi.Iteration), C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.uba_ref_time_iteration.restype = C.c_double
        L.uba_ref_time_iteration.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp, dp, ip, ip, ip, C.POINTER(capi.Calib),
                                             C.POINTER(capi.Config), C.c_int, C.c_int, dp]
        fp = C.POINTER(C.c_float); vp = C.POINTER(capi.VoParams)
        L.uba_ref_vo_project3d.argtypes = [vp, C.c_int, fp, dp]
        L.uba_ref_vo_linearize.argtypes = [vp, C.c_int, fp, dp, C.c_int, ip, dp, dp, dp, dp]
        L.uba_ref_vo_optimize.restype = C.c_int
        L.uba_ref_vo_optimize.argtypes = [vp, C.c_int, fp, dp, C.c_int, ip, dp, ip, ip]
        L.uba_ref_vo_inliers.restype = C.c_int
        L.uba_ref_vo_inliers.argtypes = [vp, C.c_int, fp, dp, ip]
        L.uba_ref_vo_ransac.restype = C.c_int
        L.uba_ref_vo_ransac.argtypes = [vp, C.c_int, fp, dp, C.c_int, ip, ip, ip, dp]
        L.uba_ref_max_threads.restype = C.c_int
        L.uba_ref_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def residual(M, calib, cam6, pt3, obs, cam_id=0):
    cam6 = capi.as_f64(cam6); pt3 = capi.as_f64(pt3); obs = capi.as_f64(obs)
    r = np.zeros(M); Jc = np.zeros((M, 6)); Jp = np.zeros((M, 3))
    rc = lib().uba_ref_residual(M, C.byref(calib), capi.dptr(cam6), capi.dptr(pt3), capi.dptr(obs), cam_id, capi.dptr(r), capi.dptr(Jc), capi.dptr(Jp))
    assert rc == 0
    return r, Jc, Jp


def tables(n_cams, n_pts, cam_idx, pt_idx, fixed_frames):
    cam_idx = capi.as_i32(cam_idx); pt_idx = capi.as_i32(pt_idx)
    n_obs = len(cam_idx)
    obs_order = np.zeros(n_obs, np.int32); pt_off = np.zeros(n_pts + 1, np.int64); pt_order = np.zeros(n_pts, np.int32)
    free_cam = np.zeros(n_cams, np.int32)
    rc = lib().uba_ref_tables(n_cams, n_pts, n_obs, capi.i32ptr(cam_idx), capi.i32ptr(pt_idx), fixed_frames, capi.i32ptr(obs_order),
                              capi.i64ptr(pt_off), capi.i32ptr(pt_order), capi.i32ptr(free_cam))
    assert rc == 0
    return {"obs_order": obs_order, "pt_obs_off": pt_off, "pt_order": pt_order, "free_cam": free_cam}


def linearize(win, cfg, fixed_frames, radius, cams=None, pts=None, jacobi_scale=None):
    """Returns dict of arrays shaped like capi.Handle.linearize plus step/model_cost_change/jacobi_scale."""
    cams = capi.as_f64(win.cams_init if cams is None else cams); pts = capi.as_f64(win.pts_init if pts is None else pts)
    M, nc, npt, no = win.M, win.n_cams, win.n_pts, win.n_obs
    fc = tables(nc, npt, win.cam_idx, win.pt_idx, fixed_frames)["free_cam"]
    n = 6 * int((fc >= 0).sum())
    arrays = {"residuals": np.zeros((no, M)), "weights": np.zeros(no), "cost": np.zeros(1), "grad_cams": np.zeros((nc, 6)),
              "grad_pts": np.zeros((npt, 3)), "B": np.zeros((nc, 6, 6)), "C": np.zeros((npt, 3, 3)), "W": np.zeros((no, 6, 3)),
              "S": np.zeros((n, n)), "rhs": np.zeros(n), "lm_diag_cams": np.zeros((nc, 6)), "lm_diag_pts": np.zeros((npt, 3))}
    out = capi.LinearizationOut()
    for k, a in arrays.items():
        setattr(out, k, capi.dptr(a))
    js_out = np.zeros(6 * nc + 3 * npt); step_c = np.zeros((nc, 6)); step_p = np.zeros((npt, 3)); mcc = np.zeros(1)
    js_in = None if jacobi_scale is None else capi.as_f64(jacobi_scale)
    rc = lib().uba_ref_linearize(M, nc, npt, no, capi.dptr(cams), capi.dptr(pts), capi.dptr(win.feats), capi.i32ptr(win.cam_idx),
                                 capi.i32ptr(win.pt_idx), capi.i32ptr(win.cam_id), C.byref(win.calib), C.byref(cfg), fixed_frames,
                                 radius, capi.dptr(js_in), C.byref(out), capi.dptr(js_out), capi.dptr(step_c), capi.dptr(step_p), capi.dptr(mcc))
    arrays.update(jacobi_scale=js_out, step_cams=step_c, step_pts=step_p, model_cost_change=mcc[0], rc=rc)
    return arrays


def cost(win, cfg, cams, pts):
    cams = capi.as_f64(cams); pts = capi.as_f64(pts)
    c = np.zeros(1)
    rc = lib().uba_ref_cost(win.M, win.n_cams, win.n_pts, win.n_obs, capi.dptr(cams), capi.dptr(pts), capi.dptr(win.feats),
                            capi.i32ptr(win.cam_idx), capi.i32ptr(win.pt_idx), capi.i32ptr(win.cam_id), C.byref(win.calib), C.byref(cfg), capi.dptr(c))
    assert rc == 0
    return c[0]


def optimise(win, cfg, fixed_frames, max_records=256):
    cams = win.cams_init.copy(); pts = win.pts_init.copy()
    s = capi.Summary(); recs = (capi.Iteration * max_records)(); n = C.c_int(0); arm = C.c_int(0)
    rc = lib().uba_ref_optimise(win.M, win.n_cams, win.n_pts, win.n_obs, capi.dptr(cams), capi.dptr(pts), capi.dptr(win.feats),
                                capi.i32ptr(win.cam_idx), capi.i32ptr(win.pt_idx), capi.i32ptr(win.cam_id), C.byref(win.calib), C.byref(cfg),
                                fixed_frames, C.byref(s), recs, max_records, C.byref(n), C.byref(arm))
    its = [{f: getattr(recs[i], f) for f, _ in capi.Iteration._fields_ if f != "pad_"} for i in range(min(n.value, max_records))]
    return dict(rc=rc, cams=cams, pts=pts, summary=s.as_dict(), iterations=its, armijo_violations=arm.value)


def time_iteration(win, cfg, fixed_frames, repeats=1):
    lin = np.zeros(1)
    t = lib().uba_ref_time_iteration(win.M, win.n_cams, win.n_pts, win.n_obs, capi.dptr(win.cams_init), capi.dptr(win.pts_init),
                                     capi.dptr(win.feats), capi.i32ptr(win.cam_idx), capi.i32ptr(win.pt_idx), capi.i32ptr(win.cam_id),
                                     C.byref(win.calib), C.byref(cfg), fixed_frames, repeats, capi.dptr(lin))
    return t, lin[0]


# ---- pose-only (stereo visual odometry) oracle: oracle/uba_vo_oracle.cpp ----
def _q(quads):
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 8)
    return q, q.ctypes.data_as(C.POINTER(C.c_float))


def vo_project3d(params, quads):
    q, qp = _q(quads); out = np.zeros((len(q), 4))
    lib().uba_ref_vo_project3d(C.byref(params), len(q), qp, capi.dptr(out))
    return out


def vo_linearize(params, quads, state, selection):
    q, qp = _q(quads); state = capi.as_f64(state); sel = capi.as_i32(selection); n = len(sel)
    A = np.zeros((6, 6)); B = np.zeros(6); res = np.zeros((n, 4)); J = np.zeros((6, 4 * n))
    lib().uba_ref_vo_linearize(C.byref(params), len(q), qp, capi.dptr(state), n, capi.i32ptr(sel), capi.dptr(A), capi.dptr(B), capi.dptr(res), capi.dptr(J))
    return dict(A=A, B=B, res=res, J=J)


def vo_optimize(params, quads, init, selection):
    q, qp = _q(quads); init = capi.as_f64(init); sel = capi.as_i32(selection)
    out = np.zeros(6); it = C.c_int32(0); st = C.c_int32(0)
    ok = lib().uba_ref_vo_optimize(C.byref(params), len(q), qp, capi.dptr(init), len(sel), capi.i32ptr(sel), capi.dptr(out), C.byref(it), C.byref(st))
    return bool(ok), out, it.value, st.value


def vo_inliers(params, quads, state):
    q, qp = _q(quads); state = capi.as_f64(state); idx = np.zeros(len(q), np.int32)
    n = lib().uba_ref_vo_inliers(C.byref(params), len(q), qp, capi.dptr(state), capi.i32ptr(idx))
    return idx[:n].copy()


def vo_ransac(params, quads, init, triples):
    q, qp = _q(quads); init = capi.as_f64(init); tr = capi.as_i32(triples).reshape(-1, 3); nh = len(tr)
    cnt = np.zeros(nh, np.int32); ok = np.zeros(nh, np.int32); st = np.zeros((nh, 6))
    best = lib().uba_ref_vo_ransac(C.byref(params), len(q), qp, capi.dptr(init), nh, capi.i32ptr(tr), capi.i32ptr(cnt), capi.i32ptr(ok), capi.dptr(st))
    return dict(best=best, counts=cnt, ok=ok, states=st)
