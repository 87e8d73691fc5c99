"""Parity of the CUDA path (through the C ABI) against the CPU oracle on a B200.

Tolerances are north_star's: residual vectors and normal-equation blocks <= 1e-9 relative, final
poses / points after a fixed number of LM iterations <= 1e-6 relative, index tables bit-exact.
"""
from pathlib import Path

import numpy as np
import pytest

from uasl_motion_estimation_b200 import capi, synth

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

BLOCK_TOL = 1e-9
STATE_TOL = 1e-6
BLOCKS = ("residuals", "weights", "cost", "grad_cams", "grad_pts", "B", "C", "W", "S", "rhs", "lm_diag_cams", "lm_diag_pts")


def rel(a, b):
    a = np.asarray(a, float).reshape(-1); b = np.asarray(b, float).reshape(-1)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def make(lib, name, scale, M=4, linearizer=0, **cfgkw):
    win = synth.config_window(name, scale=scale, lib=lib, M=M)
    cfg = capi.default_config(lib, loss_kind=synth.CONFIGS[name]["loss"], linearizer=linearizer, **cfgkw)
    h = capi.Handle(cfg, lib=lib)
    h.set_problem(M, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    return win, cfg, h


CASES = [("c1", 1.0, 4), ("c2", 0.1, 4), ("c4", 0.01, 4), ("c5", 0.02, 4), ("c1", 0.25, 2), ("c2", 0.05, 2)]


@pytest.mark.parametrize("linearizer", [0, 1, 2])
@pytest.mark.parametrize("name,scale,M", CASES)
def test_linearization_blocks(gpu_lib, oracle, name, scale, M, linearizer):
    win, cfg, h = make(gpu_lib, name, scale, M, linearizer)
    t = h.tables(2); r = oracle.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, 2)
    for k in t:
        assert np.array_equal(t[k], r[k]), k
    g = h.linearize(2, 1e4); r = oracle.linearize(win, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], r[k]) < BLOCK_TOL, (k, rel(g[k], r[k]))
    assert h.timing()["kernel_launches"] > 0


@pytest.mark.parametrize("linearizer", [0, 1, 2])
@pytest.mark.parametrize("name,scale,M,iters", [("c1", 1.0, 4, 10), ("c2", 0.1, 4, 10), ("c4", 0.01, 4, 10), ("c5", 0.02, 4, 20), ("c1", 0.25, 2, 8)])
def test_fixed_iteration_trajectory(gpu_lib, oracle, name, scale, M, iters, linearizer):
    win, cfg, h = make(gpu_lib, name, scale, M, linearizer, fixed_iterations=iters)
    rc, sums = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    assert rc == 0 and o["rc"] == 0
    gi = h.iterations(0)
    assert [a["accepted"] for a in gi] == [b["accepted"] for b in o["iterations"]]
    for a, b in zip(gi, o["iterations"]):
        assert a["cost"] == pytest.approx(b["cost"], rel=1e-7)
        assert a["model_cost_change"] == pytest.approx(b["model_cost_change"], rel=1e-5, abs=1e-12)
    assert rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL
    assert sums[0].final_cost == pytest.approx(o["summary"]["final_cost"], rel=1e-8)


def test_reference_termination_rules(gpu_lib, oracle):
    win, cfg, h = make(gpu_lib, "c1", 1.0, max_solver_time_s=0.0)
    rc, sums = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    assert sums[0].termination == o["summary"]["termination"] == 1
    assert sums[0].iterations == o["summary"]["iterations"]
    assert rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL


_SOLVER_ENVS = {
    "cluster2": {},                                          # default: two-sided band Cholesky on a 2-CTA cluster
    "lookahead": {"UBA_BAND_C2": "0"},                       # one CTA, panel-warp lookahead
    "pipelined": {"UBA_PIPE_SOLVE": "1"},                    # the cluster solver running under the lineariser (experimental)
    "bcr8": {"UBA_BAND_BCR": "8"},                           # block cyclic reduction on a cluster of 8 CTAs (experimental)
    "bcr16": {"UBA_BAND_BCR": "16"},                         # ... of 16 CTAs (non-portable cluster size)
    "bcr3": {"UBA_BAND_BCR": "3"},                           # ... of 3 (uneven deal of the eliminations)
}


@pytest.mark.parametrize("variant", sorted(_SOLVER_ENVS))
def test_large_window_band_solver_variants(variant):
    """200 keyframes -> reduced system 1188 x 1188, half-bandwidth 29.  Every band-Cholesky variant must reproduce the
    oracle trajectory.  The switches are read once per process, so each case runs in a subprocess."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import oracle_binding as ob
        from uasl_motion_estimation_b200 import capi, synth
        for scale, nfix in ((0.05, 2), (0.021, 1)):
            win = synth.config_window("c4", scale=scale)
            cfg = capi.default_config(fixed_iterations=3)
            h = capi.Handle(cfg); h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
            rc, _ = h.optimise(nfix); o = ob.optimise(win, cfg, nfix)
            rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
            assert [a["accepted"] for a in h.iterations(0)] == [b["accepted"] for b in o["iterations"]]
            assert rc == 0 and rel(h.cameras(), o["cams"]) < 1e-6 and rel(h.points(), o["pts"]) < 1e-6, (scale, rel(h.cameras(), o["cams"]))
        print("OK")
    """) % (str(ROOT), str(ROOT / "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **_SOLVER_ENVS[variant]))
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout + r.stderr)[-2000:]


@pytest.mark.parametrize("solver", [0, 1])
def test_large_window_solver_paths(gpu_lib, oracle, solver):
    """Banded Cholesky (0, default) against the blocked dense Cholesky (1) on the same 1188 x 1188 system."""
    win, cfg, h = make(gpu_lib, "c4", 0.05, fixed_iterations=3, solver=solver)
    assert 6 * int((h.tables(2)["free_cam"] >= 0).sum()) > 160
    rc, sums = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    assert rc == 0
    assert [a["accepted"] for a in h.iterations(0)] == [b["accepted"] for b in o["iterations"]]
    assert rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL


def test_batch_of_windows(gpu_lib, oracle):
    wins = [synth.config_window("c1", window=i, scale=0.05 + 0.01 * i, lib=gpu_lib) for i in range(6)]
    cfg = capi.default_config(gpu_lib, max_solver_time_s=0.0)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_batch(**synth.concat_windows(wins))
    rc, sums = h.optimise(2)
    assert rc == 0
    cams = h.cameras(); pts = h.points()
    c0 = p0 = 0
    for w, win in enumerate(wins):
        o = oracle.optimise(win, cfg, 2)
        assert sums[w].iterations == o["summary"]["iterations"] and sums[w].termination == o["summary"]["termination"]
        assert rel(cams[c0:c0 + win.n_cams], o["cams"]) < STATE_TOL
        assert rel(pts[p0:p0 + win.n_pts], o["pts"]) < STATE_TOL
        c0 += win.n_cams; p0 += win.n_pts


def test_shuffled_observations_ragged_and_empty_points(gpu_lib, oracle):
    win = synth.config_window("c2", scale=0.05, lib=gpu_lib)
    rng = np.random.default_rng(0)
    perm = rng.permutation(win.n_obs)
    keep = perm[win.pt_idx[perm] % 7 != 3]
    sub = synth.Window(4, win.cams_gt, win.cams_init, win.pts_gt, win.pts_init, np.ascontiguousarray(win.feats[keep]),
                       np.ascontiguousarray(win.cam_idx[keep]), np.ascontiguousarray(win.pt_idx[keep]),
                       np.ascontiguousarray(win.cam_id[keep]), 2, win.calib)
    cfg = capi.default_config(gpu_lib, fixed_iterations=5)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_problem(4, sub.cams_init, sub.pts_init, sub.feats, sub.cam_idx, sub.pt_idx, sub.cam_id, sub.calib)
    g = h.linearize(2, 1e4); r = oracle.linearize(sub, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], r[k]) < BLOCK_TOL, k
    rc, _ = h.optimise(2)
    o = oracle.optimise(sub, cfg, 2)
    assert rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL
    untouched = np.flatnonzero(np.bincount(sub.pt_idx, minlength=sub.n_pts) == 0)
    np.testing.assert_array_equal(h.points()[untouched], sub.pts_init[untouched])


def test_errors_and_infeasible_start(gpu_lib):
    win = synth.config_window("c1", scale=0.05, lib=gpu_lib)
    h = capi.Handle(capi.default_config(gpu_lib), lib=gpu_lib)
    bad = win.cam_idx.copy(); bad[5] = -1
    with pytest.raises(capi.UbaError) as e:
        h.set_problem(4, win.cams_init, win.pts_init, win.feats, bad, win.pt_idx, None, win.calib)
    assert e.value.code == capi.UBA_ERR_INVALID_ARGUMENT
    pts = win.pts_init.copy(); pts[3, 2] = 1e9
    h.set_problem(4, win.cams_init, pts, win.feats, win.cam_idx, win.pt_idx, None, win.calib)
    rc, sums = h.optimise(2, check=False)
    assert rc == capi.UBA_ERR_INFEASIBLE and sums[0].usable == 0
    np.testing.assert_array_equal(h.points(), pts)


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_full_size_properties(gpu_lib, oracle, name):
    """BASELINE.json's full sizes: size-independent properties instead of a full oracle LM run."""
    win, cfg, h = make(gpu_lib, name, 1.0, fixed_iterations=synth.CONFIGS[name]["iters"])
    g = h.linearize(2, 1e4, want=("cost", "S", "rhs", "grad_cams", "residuals"))
    # cost and residuals against the oracle's residual-only evaluation (cheap even at 1M observations)
    assert g["cost"][0] == pytest.approx(oracle.cost(win, cfg, win.cams_init, win.pts_init), rel=1e-11)
    S = g["S"]
    assert np.abs(S - S.T).max() <= 1e-12 * np.abs(S).max()
    np.linalg.cholesky(S)  # positive definite
    # linearisation is idempotent (atomics may reorder sums: tolerance, not bitwise)
    g2 = h.linearize(2, 1e4, want=("S", "rhs"))
    assert rel(g2["S"], S) < 1e-12 and rel(g2["rhs"], g["rhs"]) < 1e-12
    # directional derivative of the oracle's cost == gradient . direction
    rng = np.random.default_rng(3)
    d = rng.normal(size=(win.n_cams, 6)) * np.array([1e-3] * 3 + [1e-4] * 3); d[:2] = 0
    eps = 1e-4
    fd = (oracle.cost(win, cfg, win.cams_init + eps * d, win.pts_init) - oracle.cost(win, cfg, win.cams_init - eps * d, win.pts_init)) / (2 * eps)
    if cfg.loss_kind == capi.LOSS_HUBER:  # Huber's kink makes the FD check only approximate
        assert np.sum(g["grad_cams"] * d) == pytest.approx(fd, rel=1e-3)
    rc, sums = h.optimise(2)
    its = h.iterations(0)
    assert rc == 0 and sums[0].final_cost < (0.2 if name == "c4" else 0.9) * sums[0].initial_cost
    costs = [it["cost"] for it in its]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(costs, costs[1:]))  # monotonic: only accepted steps move x
    assert sums[0].final_cost == pytest.approx(oracle.cost(win, cfg, h.cameras(), h.points()), rel=1e-10)


@pytest.mark.parametrize("name,scale", [("c1", 0.25), ("c4", 0.02)])
def test_pose_covariances(gpu_lib, oracle, name, scale):
    """extract_covariance (BundleAdjuster.h:478-528): 6x6 blocks of (J^T J)^-1 with the points marginalised ==
    diagonal blocks of the inverse of the undamped reduced camera matrix the oracle builds at the final iterate."""
    win, cfg, h = make(gpu_lib, name, scale, fixed_iterations=6, compute_covariance=1)
    rc, sums = h.optimise(2)
    assert rc == 0
    cov = h.pose_covariances()
    cams, pts = h.cameras(), h.points()
    L = oracle.linearize(win, cfg, 2, -1.0, cams=cams, pts=pts, jacobi_scale=np.ones(6 * win.n_cams + 3 * win.n_pts))
    Sinv = np.linalg.inv(L["S"])
    fc = h.tables(2)["free_cam"]
    for c in range(win.n_cams):
        if fc[c] < 0:
            assert not cov[c].any()
        else:
            ref = Sinv[6 * fc[c]:6 * fc[c] + 6, 6 * fc[c]:6 * fc[c] + 6]
            assert np.abs(cov[c] - ref).max() <= 1e-7 * np.abs(ref).max(), c
            assert np.allclose(cov[c], cov[c].T, rtol=1e-10, atol=0) and np.all(np.linalg.eigvalsh(cov[c]) > 0)
    # results of the optimisation itself are unaffected by the extra pass
    o = oracle.optimise(win, cfg, 2)
    assert rel(cams, o["cams"]) < STATE_TOL and rel(pts, o["pts"]) < STATE_TOL


def test_long_tracks_fall_back_to_the_generic_lineariser(gpu_lib, oracle):
    """30 keyframes: full tracks (28 free cameras > the tile limit of 21) mixed with short ones -> the tiled kernel
    and the generic kernel both accumulate into the same reduced system."""
    long_w = synth.generate(30, 60, 30, 30, full_tracks=1, seed=11, lib=gpu_lib)
    short_w = synth.generate(30, 400, 3, 9, seed=12, lib=gpu_lib)
    cat = lambda a, b: np.ascontiguousarray(np.concatenate([a, b]))
    win = synth.Window(4, long_w.cams_gt, long_w.cams_init, cat(long_w.pts_gt, short_w.pts_gt), cat(long_w.pts_init, short_w.pts_init),
                       cat(long_w.feats, short_w.feats), cat(long_w.cam_idx, short_w.cam_idx),
                       cat(long_w.pt_idx, short_w.pt_idx + long_w.n_pts).astype(np.int32), cat(long_w.cam_id, short_w.cam_id), 2, long_w.calib)
    cfg = capi.default_config(gpu_lib, fixed_iterations=5)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    g = h.linearize(2, 1e4); r = oracle.linearize(win, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], r[k]) < BLOCK_TOL, (k, rel(g[k], r[k]))
    rc, _ = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    assert rc == 0 and rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL


def test_duplicate_observations_of_one_camera(gpu_lib, oracle):
    """The reference accepts two residual blocks on the same (camera, point) pair; such points bypass the tile plan."""
    win = synth.config_window("c1", scale=0.05, lib=gpu_lib)
    dup = np.flatnonzero((win.pt_idx % 5 == 0) & (win.cam_idx == 4))
    feats = np.concatenate([win.feats, win.feats[dup] + 0.25]); cam_idx = np.concatenate([win.cam_idx, win.cam_idx[dup]])
    pt_idx = np.concatenate([win.pt_idx, win.pt_idx[dup]]); cam_id = np.concatenate([win.cam_id, win.cam_id[dup]])
    w2 = synth.Window(4, win.cams_gt, win.cams_init, win.pts_gt, win.pts_init, np.ascontiguousarray(feats), cam_idx.astype(np.int32),
                      pt_idx.astype(np.int32), cam_id.astype(np.int32), 2, win.calib)
    cfg = capi.default_config(gpu_lib, fixed_iterations=4)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_problem(4, w2.cams_init, w2.pts_init, w2.feats, w2.cam_idx, w2.pt_idx, w2.cam_id, w2.calib)
    g = h.linearize(2, 1e4); r = oracle.linearize(w2, cfg, 2, 1e4)
    for k in ("residuals", "cost", "B", "C", "S", "rhs", "grad_cams", "grad_pts"):
        assert rel(g[k], r[k]) < BLOCK_TOL, (k, rel(g[k], r[k]))
    rc, _ = h.optimise(2)
    o = oracle.optimise(w2, cfg, 2)
    assert rc == 0 and rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL


def test_mono_batch_and_window_without_free_cameras(gpu_lib, oracle):
    """M = 2 batch; the last window keeps every camera fixed (points only)."""
    wins = [synth.config_window("c1", window=i, scale=0.04, lib=gpu_lib, M=2) for i in range(3)]
    cfg = capi.default_config(gpu_lib, fixed_iterations=4)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_batch(**synth.concat_windows(wins))
    rc, sums = h.optimise(2)
    assert rc == 0
    cams = h.cameras(); pts = h.points(); c0 = p0 = 0
    for win in wins:
        o = oracle.optimise(win, cfg, 2)
        assert rel(cams[c0:c0 + win.n_cams], o["cams"]) < STATE_TOL and rel(pts[p0:p0 + win.n_pts], o["pts"]) < STATE_TOL
        c0 += win.n_cams; p0 += win.n_pts
    h2 = capi.Handle(cfg, lib=gpu_lib)
    w = wins[0]
    h2.set_problem(2, w.cams_init, w.pts_init, w.feats, w.cam_idx, w.pt_idx, w.cam_id, w.calib)
    rc, _ = h2.optimise(w.n_cams)
    o = oracle.optimise(w, cfg, w.n_cams)
    np.testing.assert_array_equal(h2.cameras(), w.cams_init)
    assert rel(h2.points(), o["pts"]) < STATE_TOL


def test_iteration_graph_is_rebuilt_when_the_solver_settings_change(gpu_lib, oracle):
    """The LM iteration is replayed from a CUDA graph that bakes the device view in by value.  Re-submitting a window of
    the same shape reuses the graph; changing anything in the view (here the iteration budget, through the timing entry
    point) must not: a stale graph would stop iterating after the old budget and report a far too short time."""
    win, cfg, h = make(gpu_lib, "c2", 0.5, fixed_iterations=2)
    o = oracle.optimise(win, cfg, 2)
    for _ in range(2):   # second call replays the captured graph
        h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
        rc, sums = h.optimise(2)
        assert rc == 0 and sums[0].iterations == 2
        assert rel(h.cameras(), o["cams"]) < STATE_TOL and rel(h.points(), o["pts"]) < STATE_TOL
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    ms_iter = h.time_iteration(2, iterations=8, flush_l2=False)
    ms_lin = h.time_linearize(2, 1e4, repeats=8, flush_l2=False)
    assert ms_iter > ms_lin, (ms_iter, ms_lin)   # eight full iterations, not two plus six early exits
    # and a different window of the same shape through the same handle still gives that window's answer
    win2 = synth.config_window("c2", window=3, scale=0.5, lib=gpu_lib)
    h.set_problem(4, win2.cams_init, win2.pts_init, win2.feats, win2.cam_idx, win2.pt_idx, win2.cam_id, win2.calib)
    rc, sums = h.optimise(2)
    o2 = oracle.optimise(win2, cfg, 2)
    assert rc == 0 and rel(h.cameras(), o2["cams"]) < STATE_TOL and rel(h.points(), o2["pts"]) < STATE_TOL


@pytest.mark.parametrize("variant", ["track_order", "camera_descending", "random_points"])
def test_ingest_paths(gpu_lib, oracle, variant):
    """Canonical input (fast path: one streaming copy, identity maps) and permuted input (general path) through the device-side
    feature gather; results against the oracle, and the canonical and permuted runs against each other."""
    base = synth.config_window("c4", scale=0.05, lib=gpu_lib)
    if variant == "track_order":
        win = synth.reorder(base)
    elif variant == "camera_descending":
        win = synth.reorder(base, camera_descending=True)
    else:
        win = synth.reorder(base, point_order=np.random.default_rng(3).permutation(base.n_pts))
    cfg = capi.default_config(gpu_lib, fixed_iterations=4)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    t = h.tables(2); r = oracle.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, 2)
    for k in t:
        assert np.array_equal(t[k], r[k]), k
    g = h.linearize(2, 1e4); o = oracle.linearize(win, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], o[k]) < BLOCK_TOL, k
    rc, _ = h.optimise(2)
    oo = oracle.optimise(win, cfg, 2)
    assert rc == 0 and rel(h.cameras(), oo["cams"]) < STATE_TOL and rel(h.points(), oo["pts"]) < STATE_TOL


def test_small_window_left_looking_solver_fallback():
    """UBA_SMALL_LA=0 selects the left-looking dense solver for small windows; same trajectory as the oracle (c1 and c2 shapes,
    and the covariance pass, which asks the solver to keep its factor)."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import oracle_binding as ob
        from uasl_motion_estimation_b200 import capi, synth
        rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
        for name, scale in (("c1", 0.5), ("c2", 0.1)):
            win = synth.config_window(name, scale=scale)
            cfg = capi.default_config(fixed_iterations=4, compute_covariance=1)
            h = capi.Handle(cfg); h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
            rc, _ = h.optimise(2); o = ob.optimise(win, cfg, 2)
            assert rc == 0 and rel(h.cameras(), o["cams"]) < 1e-6 and rel(h.points(), o["pts"]) < 1e-6
            cov = h.pose_covariances(); cams, pts = h.cameras(), h.points()
            L = ob.linearize(win, cfg, 2, -1.0, cams=cams, pts=pts, jacobi_scale=np.ones(6 * win.n_cams + 3 * win.n_pts))
            Sinv = np.linalg.inv(L["S"]); fc = h.tables(2)["free_cam"]
            for c in range(win.n_cams):
                if fc[c] >= 0:
                    ref = Sinv[6 * fc[c]:6 * fc[c] + 6, 6 * fc[c]:6 * fc[c] + 6]
                    assert np.abs(cov[c] - ref).max() <= 1e-7 * np.abs(ref).max(), c
        print("OK")
    """) % (str(ROOT), str(ROOT / "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, UBA_SMALL_LA="0"))
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout + r.stderr)[-2000:]


def test_features_that_are_not_floats_take_the_double_upload(gpu_lib, oracle):
    """The reference's features are float detections widened to double, and such rows travel to the GPU as float32 (exactly).
    Rows with any value that is not a float must take the double path and still match the oracle on the same doubles."""
    base = synth.config_window("c2", scale=0.01, lib=gpu_lib)
    feats = base.feats + 1e-7 * np.sin(np.arange(base.feats.size)).reshape(base.feats.shape)   # no longer float32-representable
    assert (feats.astype(np.float32).astype(np.float64) != feats).any()
    win = synth.Window(4, base.cams_gt, base.cams_init, base.pts_gt, base.pts_init, np.ascontiguousarray(feats), base.cam_idx, base.pt_idx,
                       base.cam_id, 2, base.calib)
    cfg = capi.default_config(gpu_lib, fixed_iterations=3)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    g = h.linearize(2, 1e4); o = oracle.linearize(win, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], o[k]) < BLOCK_TOL, k
    # the residuals see the perturbation: they differ from the float-rounded window's by more than the parity tolerance
    hb = capi.Handle(cfg, lib=gpu_lib)
    hb.set_problem(4, base.cams_init, base.pts_init, base.feats, base.cam_idx, base.pt_idx, base.cam_id, base.calib)
    assert rel(hb.linearize(2, 1e4)["residuals"], g["residuals"]) > 1e-9


# ---- BASELINE.json's full sizes against the oracle itself (the oracle does a c4 iteration in ~0.3 s) ----------------------
@pytest.mark.parametrize("name", ["c2", "c4", "c5"])
def test_full_size_blocks_and_trajectory_match_the_oracle(gpu_lib, oracle, name):
    """Scale 1.0: every normal-equation block <= 1e-9 and the K-iteration trajectory (K = 10; c5: 20) <= 1e-6 with the
    same accept / reject sequence — on the reduced systems the bench times (c4: 1188 x 1188, c5: 588 x 588, c2: 108 x 108)."""
    K = synth.CONFIGS[name]["iters"]
    win, cfg, h = make(gpu_lib, name, 1.0, fixed_iterations=K)
    blocks = ("residuals", "cost", "grad_cams", "grad_pts", "B", "C", "S", "rhs")
    g = h.linearize(2, 1e4, want=blocks)
    r = oracle.linearize(win, cfg, 2, 1e4)
    for k in blocks:
        assert rel(g[k], r[k]) < BLOCK_TOL, (k, rel(g[k], r[k]))
    rc, sums = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    assert rc == 0 and sums[0].iterations == K == o["summary"]["iterations"]
    assert [a["accepted"] for a in h.iterations(0)] == [b["accepted"] for b in o["iterations"]]
    assert rel(h.cameras(), o["cams"]) < STATE_TOL
    if name == "c5":
        # 30 % outliers, Cauchy, 20 iterations: a few dozen of the 100 000 points are driven kilometres away and are barely
        # constrained.  The cloud as a whole meets 1e-6 (Frobenius); the worst single coordinate is
        # allowed 1e-5 of the cloud's extent, and all but a handful of points must meet 1e-6 of their own size one by one.
        dp = np.abs(h.points() - o["pts"])
        assert np.linalg.norm(dp) / np.linalg.norm(o["pts"]) < STATE_TOL
        assert rel(h.points(), o["pts"]) < 1e-5
        per_pt = dp.max(axis=1) / np.maximum(np.abs(o["pts"]).max(axis=1), 1.0)
        assert (per_pt > 1e-6).sum() <= 50, int((per_pt > 1e-6).sum())
    else:
        assert rel(h.points(), o["pts"]) < STATE_TOL
    assert abs(sums[0].final_cost - o["summary"]["final_cost"]) <= 1e-9 * o["summary"]["final_cost"]


def test_batch_of_64_full_size_windows_matches_the_oracle(gpu_lib, oracle):
    """A slice of c3: 64 independent 10-frame windows (2 000 points, 20 000 observations each) in one handle."""
    wins = [synth.config_window("c3", window=i, lib=gpu_lib) for i in range(64)]
    cfg = capi.default_config(gpu_lib, fixed_iterations=10)
    h = capi.Handle(cfg, lib=gpu_lib)
    h.set_batch(**synth.concat_windows(wins))
    rc, sums = h.optimise(2)
    assert rc == 0
    cams = h.cameras(); pts = h.points()
    c0 = p0 = 0
    for w, win in enumerate(wins):
        if w % 8 == 0:       # every eighth window through the oracle (0.1 s each)
            o = oracle.optimise(win, cfg, 2)
            assert rel(cams[c0:c0 + win.n_cams], o["cams"]) < STATE_TOL, w
            assert rel(pts[p0:p0 + win.n_pts], o["pts"]) < STATE_TOL, w
            assert abs(sums[w].final_cost - o["summary"]["final_cost"]) <= 1e-9 * o["summary"]["final_cost"]
        c0 += win.n_cams; p0 += win.n_pts


def test_repeated_runs_agree(gpu_lib):
    """fp64 atomics reorder sums between runs: poses must agree to rounding level, points to far better than the parity gate
    (c5's barely constrained outlier points are the worst case: DESIGN.md section 5)."""
    win, cfg, h = make(gpu_lib, "c5", 0.2, fixed_iterations=20)
    h.optimise(2); c1, p1 = h.cameras(), h.points()
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    h.optimise(2); c2, p2 = h.cameras(), h.points()
    assert rel(c1, c2) < 1e-11 and rel(p1, p2) < 1e-7
