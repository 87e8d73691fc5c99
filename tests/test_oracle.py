"""The CPU oracle against independent evidence: 40-digit golden vectors written straight from the
reference's functors (tests/golden/make_golden.py), OpenCV's Rodrigues, finite differences, scipy.
The reference ships no tests or golden vectors of its own (parity is otherwise unpinned)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from uasl_motion_estimation_b200 import capi

GOLD = json.loads((Path(__file__).parent / "golden" / "residual_golden.json").read_text())


def _calib():
    k = capi.Calib()
    g = GOLD["calib"]
    k.fx0, k.fy0, k.cx0, k.cy0 = g["fx0"], g["fy0"], g["cx0"], g["cy0"]
    k.fx1, k.fy1, k.cx1, k.cy1 = g["fx1"], g["fy0"], g["cx1"], g["cy0"]
    k.feat_var, k.baseline = g["feat_var"], g["baseline"]
    return k


@pytest.mark.parametrize("i", range(len(GOLD["cases"])))
def test_residual_and_jacobian_vs_mpmath_golden(oracle, i):
    c = GOLD["cases"][i]
    r, Jc, Jp = oracle.residual(c["M"], _calib(), c["cam"], c["X"], c["obs"], c["cam_id"])
    J = np.hstack([Jc, Jp])
    np.testing.assert_allclose(r, c["r"], rtol=1e-12, atol=1e-10)  # residuals are O(1e3) px/sigma
    np.testing.assert_allclose(J, np.array(c["J"]), rtol=1e-10, atol=1e-9)


def test_rotation_matches_opencv_rodrigues(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(50):
        r = rng.normal(size=3) * rng.choice([1e-4, 0.1, 2.0])
        p = rng.normal(size=3) * 10
        out = np.zeros(3)
        oracle.lib().uba_ref_rotate(capi.dptr(r), capi.dptr(p), capi.dptr(out))
        R, _ = cv2.Rodrigues(r)
        np.testing.assert_allclose(out, R @ p, rtol=1e-12, atol=1e-12)


def test_small_angle_branch_is_first_order(oracle):
    r = np.array([1e-9, -2e-9, 0.5e-9]); p = np.array([1.0, 2.0, 3.0]); out = np.zeros(3)
    oracle.lib().uba_ref_rotate(capi.dptr(r), capi.dptr(p), capi.dptr(out))
    np.testing.assert_array_equal(out, p + np.cross(r, p))


@pytest.mark.parametrize("kind,fn", [(capi.LOSS_HUBER, lambda s: (s, 1.0) if s <= 1 else (2 * np.sqrt(s) - 1, 1 / np.sqrt(s))),
                                     (capi.LOSS_CAUCHY, lambda s: (np.log1p(s), 1 / (1 + s))),
                                     (capi.LOSS_TRIVIAL, lambda s: (s, 1.0))])
def test_loss_functions(oracle, kind, fn):
    for s in [0.0, 0.3, 1.0, 1.0000001, 7.5, 1e6]:
        rho = np.zeros(3)
        oracle.lib().uba_ref_loss(kind, 1.0, s, capi.dptr(rho))
        e0, e1 = fn(s)
        assert rho[0] == pytest.approx(e0, rel=1e-14) and rho[1] == pytest.approx(e1, rel=1e-14)
        assert rho[2] <= 0.0  # the corrector's plain sqrt(rho') scaling is valid


def test_point_bounds_follow_the_reference_formulas(oracle):
    k = _calib(); lo = np.zeros(3); hi = np.zeros(3)
    oracle.lib().uba_ref_point_bounds(C.byref(k), 4, capi.dptr(lo), capi.dptr(hi))
    zmax = k.fx0 * k.baseline / 0.1; zmin = k.fx0 * k.baseline / (2 * k.cx0)  # BundleAdjuster.h:442-443
    np.testing.assert_allclose(hi, [zmax / k.fx0 * k.cx0, zmax / k.fy0 * k.cy0, zmax])
    np.testing.assert_allclose(lo, [-zmax / k.fx0 * k.cx0, -zmax / k.fy0 * k.cy0, zmin])
    k.baseline = 0.0  # mono: zero baseline becomes 0.5 (:389-390)
    oracle.lib().uba_ref_point_bounds(C.byref(k), 2, capi.dptr(lo), capi.dptr(hi))
    assert hi[2] == pytest.approx(k.fx0 * 0.5 / 0.1)


def test_quaternion_maps_round_trip(oracle):
    rng = np.random.default_rng(2)
    for _ in range(20):
        r = rng.normal(size=3) * 0.4
        q = np.zeros(4); r2 = np.zeros(3)
        oracle.lib().uba_ref_exp_map_quat(capi.dptr(r), capi.dptr(q))
        oracle.lib().uba_ref_log_map_quat(capi.dptr(q), capi.dptr(r2))
        np.testing.assert_allclose(r2, r, rtol=1e-12, atol=1e-14)
        assert np.linalg.norm(q) == pytest.approx(1.0)
    q = np.array([1.0, 0, 0, 0]); r = np.ones(3)
    oracle.lib().uba_ref_log_map_quat(capi.dptr(q), capi.dptr(r))  # identity -> r = 0 (small-angle branch)
    np.testing.assert_array_equal(r, 0)


def test_schur_system_matches_dense_normal_equations(oracle, emu_lib):
    """S and rhs of the oracle == Schur complement of a dense numpy J^T J built from its own blocks."""
    from uasl_motion_estimation_b200 import synth
    win = synth.config_window("c2", scale=0.002, lib=emu_lib)
    cfg = capi.default_config(emu_lib)
    L = oracle.linearize(win, cfg, 2, 1e4)
    nc, npt, M = win.n_cams, win.n_pts, win.M
    fc = oracle.tables(nc, npt, win.cam_idx, win.pt_idx, 2)["free_cam"]
    nf = int((fc >= 0).sum())
    # rebuild J~ from residual-level autodiff
    J = np.zeros((win.n_obs * M, 6 * nf + 3 * npt)); r = np.zeros(win.n_obs * M)
    for o in range(win.n_obs):
        rr, Jc, Jp = oracle.residual(M, win.calib, win.cams_init[win.cam_idx[o]], win.pts_init[win.pt_idx[o]], win.feats[o], int(win.cam_id[o]))
        w = L["weights"][o]
        if fc[win.cam_idx[o]] >= 0:
            J[o * M:(o + 1) * M, 6 * fc[win.cam_idx[o]]:6 * fc[win.cam_idx[o]] + 6] = w * Jc
        J[o * M:(o + 1) * M, 6 * nf + 3 * win.pt_idx[o]:6 * nf + 3 * win.pt_idx[o] + 3] = w * Jp
        r[o * M:(o + 1) * M] = w * rr
    H = J.T @ J; g = J.T @ r
    lam = np.concatenate([L["lm_diag_cams"][fc >= 0].reshape(-1), L["lm_diag_pts"].reshape(-1)])
    H = H + np.diag(lam)
    n = 6 * nf
    active = np.repeat(np.bincount(win.pt_idx, minlength=npt) > 0, 3)
    Hpp = H[n:, n:][np.ix_(active, active)]; Hcp = H[:n, n:][:, active]
    S = H[:n, :n] - Hcp @ np.linalg.solve(Hpp, Hcp.T)
    rhs = g[:n] - Hcp @ np.linalg.solve(Hpp, g[n:][active])
    np.testing.assert_allclose(L["S"], S, rtol=1e-9, atol=1e-9 * np.abs(S).max())
    np.testing.assert_allclose(L["rhs"], rhs, rtol=1e-9, atol=1e-9 * np.abs(rhs).max())
    # and the LM step solves the damped normal equations
    y = np.linalg.solve(H[np.ix_(np.r_[np.arange(n), n + np.flatnonzero(active)], np.r_[np.arange(n), n + np.flatnonzero(active)])],
                        np.r_[g[:n], g[n:][active]])
    step_c = L["step_cams"][fc >= 0].reshape(-1)
    np.testing.assert_allclose(step_c, -y[:n], rtol=1e-7, atol=1e-9 * np.abs(y).max())


def test_oracle_optimise_reduces_cost_and_agrees_with_scipy(oracle, emu_lib):
    from scipy.optimize import least_squares

    from uasl_motion_estimation_b200 import synth
    win = synth.config_window("c1", scale=0.01, lib=emu_lib)
    cfg = capi.default_config(emu_lib, loss_kind=capi.LOSS_TRIVIAL, max_iterations=50, function_tolerance=1e-12, max_solver_time_s=0.0)
    o = oracle.optimise(win, cfg, 2)
    assert o["summary"]["usable"] == 1 and o["summary"]["final_cost"] < 0.1 * o["summary"]["initial_cost"]
    nc, npt = win.n_cams, win.n_pts

    def fun(x):
        cams = win.cams_init.copy(); cams[2:] = x[:6 * (nc - 2)].reshape(-1, 6); pts = x[6 * (nc - 2):].reshape(-1, 3)
        out = np.zeros((win.n_obs, 4))
        for i in range(win.n_obs):
            out[i] = oracle.residual(4, win.calib, cams[win.cam_idx[i]], pts[win.pt_idx[i]], win.feats[i])[0]
        return out.reshape(-1)
    x0 = np.r_[o["cams"][2:].reshape(-1), o["pts"].reshape(-1)]
    sol = least_squares(fun, x0, method="lm", max_nfev=20)
    assert 0.5 * np.sum(sol.fun ** 2) == pytest.approx(o["summary"]["final_cost"], rel=1e-6)


def test_oracle_infeasible_start_fails(oracle, emu_lib):
    from uasl_motion_estimation_b200 import synth
    win = synth.config_window("c1", scale=0.01, lib=emu_lib)
    win.pts_init[3, 2] = 1e9  # outside Zmax
    before = win.pts_init.copy()
    o = oracle.optimise(win, capi.default_config(emu_lib), 2)
    assert o["rc"] == capi.UBA_ERR_INFEASIBLE and o["summary"]["usable"] == 0
    np.testing.assert_array_equal(o["pts"], before)


# ---- boundary maths of the product library (libuba_host.so: same sources as libuba.so) against the oracle's copy -----------
def test_library_log_exp_maps_match_the_oracle_and_clamp():
    """uba_log_map_quat / uba_exp_map_quat (core/rotation_utils.h:190-204) of the PRODUCT library: equal to the oracle's
    restatement on random rotations, inverse of each other, and clamped where the reference's acos(w > 1) gives NaN."""
    import oracle_binding as ob
    from uasl_motion_estimation_b200 import capi
    lib = capi.load_host()
    rng = np.random.default_rng(5)
    for _ in range(200):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if q[0] < 0: q = -q
        r_lib = np.zeros(3); r_ref = np.zeros(3)
        lib.uba_log_map_quat(capi.dptr(q), capi.dptr(r_lib)); ob.lib().uba_ref_log_map_quat(capi.dptr(q), capi.dptr(r_ref))
        assert np.array_equal(r_lib, r_ref)
        q_lib = np.zeros(4); q_ref = np.zeros(4)
        lib.uba_exp_map_quat(capi.dptr(r_lib), capi.dptr(q_lib)); ob.lib().uba_ref_exp_map_quat(capi.dptr(r_lib), capi.dptr(q_ref))
        assert np.array_equal(q_lib, q_ref)
        assert np.abs(q_lib - q).max() < 1e-12
    # identity and the w a hair above one case (rotation_utils.h:203 has no clamp: NaN in the reference)
    for w in (1.0, 1.0 + 2.3e-16):
        q = np.array([w, 0.0, 0.0, 0.0]); r = np.ones(3)
        lib.uba_log_map_quat(capi.dptr(q), capi.dptr(r))
        assert np.array_equal(r, np.zeros(3))
    q = np.zeros(4); lib.uba_exp_map_quat(capi.dptr(np.zeros(3)), capi.dptr(q))
    assert np.array_equal(q, np.array([1.0, 0.0, 0.0, 0.0]))
