"""Pose-only mode (stereo visual odometry core).

CPU (-m "not gpu"): the oracle restatement (oracle/uba_vo_oracle.cpp) against the reference's own StereoVisualOdometry.cpp
compiled in oracle/_ref — project3D, reproject + residuals, updateJacobian, J J^T, J r, computeInliers — and the
Gauss-Newton / Levenberg-Marquardt trajectories checked step by step against the reference's linearisation.
GPU (-m gpu): libuba's uba_vo_* entry points against the oracle."""
import numpy as np
import pytest

import ref_binding as rb
from uasl_motion_estimation_b200 import capi, synth

rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
needs_ref = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (no reference tree here)")
TRUE_MOTION_HINT = np.array([0.0, 0.0, 0.0, 0.0, 0.0, -0.8])   # the rig advances 0.8 m per keyframe


def _setup(n=400, outliers=0.0, **over):
    quads, out = synth.vo_quads(n, outliers)
    calib = capi.default_calib()
    return quads, out, synth.vo_params(calib, **over), rb.vo_params(calib) if rb.available() else None


@needs_ref
def test_oracle_triangulation_and_linearisation_equal_the_reference(oracle):
    quads, _, P, p10 = _setup()
    assert rel(oracle.vo_project3d(P, quads), rb.vo_project3d(p10, quads)) < 1e-15
    rng = np.random.default_rng(3)
    for trial in range(6):
        state = np.zeros(6) if trial == 0 else np.concatenate([rng.normal(size=3) * 0.02, rng.normal(size=3) * 0.5])
        sel = rng.choice(len(quads), size=3 if trial % 2 else 60, replace=False).astype(np.int32)
        a = oracle.vo_linearize(P, quads, state, sel); b = rb.vo_linearize(p10, quads, state, sel)
        assert rel(a["res"], b["res"]) < 1e-13 and rel(a["J"], b["J"]) < 1e-12
        assert rel(a["A"], b["A"]) < 1e-12 and rel(a["B"], b["B"]) < 1e-11
        assert np.array_equal(oracle.vo_inliers(P, quads, state), rb.vo_inliers(p10, quads, state))


@needs_ref
@pytest.mark.parametrize("method", [0, 1])
def test_oracle_trajectory_follows_the_reference_linearisation(oracle, method):
    """optimize() of the reference cannot be run to completion (its loop condition, :277, never lets it return on ordinary
    data), so the iteration is pinned step by step: running the oracle for k = 1, 2, 3... iterations, every Gauss-Newton
    step must equal solve(A_ref, B_ref) at the previous iterate, with A_ref, B_ref from the compiled reference."""
    quads, _, _, p10 = _setup()
    sel = np.arange(0, 120, dtype=np.int32)
    prev = np.zeros(6)
    for k in range(1, 6):
        P = synth.vo_params(capi.default_calib(), method=method, max_iter=k, e1=0.0, e2=0.0, e3=0.0, e4=0.0)
        ok, state, iters, stop = oracle.vo_optimize(P, quads, np.zeros(6), sel)
        assert iters == k and stop == 3 and not ok           # MAX_ITERATIONS -> false (:279-280)
        if method == 0:
            lin = rb.vo_linearize(p10, quads, prev, sel)
            step = np.linalg.solve(lin["A"], lin["B"])
            assert np.abs((state - prev) - step).max() < 1e-9 * max(1.0, np.abs(step).max())
        prev = state
    # and it converges to the motion of the rig
    P = synth.vo_params(capi.default_calib(), method=method)
    ok, state, iters, stop = oracle.vo_optimize(P, quads, np.zeros(6), sel)
    assert ok and iters < 30 and abs(state[5] + 0.8) < 0.05 and np.abs(state[:3]).max() < 0.01


def test_oracle_ransac_recovers_the_motion_under_outliers(oracle):
    quads, out, P, _ = _setup(600, 0.3)
    rng = np.random.default_rng(11)
    triples = np.stack([rng.choice(len(quads), 3, replace=False) for _ in range(60)]).astype(np.int32)
    r = oracle.vo_ransac(P, quads, np.zeros(6), triples)
    assert r["best"] >= 0 and r["counts"][r["best"]] == r["counts"].max()
    inl = oracle.vo_inliers(P, quads, r["states"][r["best"]])
    assert len(inl) == r["counts"][r["best"]] and out[inl].mean() < 0.02           # the consensus set is (almost) outlier-free
    ok, state, _, _ = oracle.vo_optimize(P, quads, np.zeros(6), inl)
    assert ok and abs(state[5] + 0.8) < 0.05


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_triangulation_linearisation_and_inliers_match_the_oracle(gpu_lib, oracle):
    quads, _, P, _ = _setup()
    with capi.Handle(capi.default_config(gpu_lib), lib=gpu_lib) as h:
        h.vo_set_matches(P, quads)
        assert rel(h.vo_points(), oracle.vo_project3d(P, quads)) < 1e-14
        rng = np.random.default_rng(5)
        for trial in range(6):
            state = np.zeros(6) if trial == 0 else np.concatenate([rng.normal(size=3) * 0.02, rng.normal(size=3) * 0.5])
            sel = rng.choice(len(quads), size=3 if trial % 2 else 200, replace=False).astype(np.int32)
            g = h.vo_linearize(state, sel); o = oracle.vo_linearize(P, quads, state, sel)
            assert rel(g["res"], o["res"]) < 1e-9 and rel(g["J"], o["J"]) < 1e-9           # north_star: blocks within 1e-9
            assert rel(g["A"], o["A"]) < 1e-9 and rel(g["B"], o["B"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1])
def test_gpu_refinement_matches_the_oracle(gpu_lib, oracle, method):
    quads, _, P, _ = _setup(method=method)
    with capi.Handle(capi.default_config(gpu_lib), lib=gpu_lib) as h:
        h.vo_set_matches(P, quads)
        for sel in (np.arange(3, dtype=np.int32) * 50, np.arange(len(quads), dtype=np.int32)):
            ok, state, iters = h.vo_refine(np.zeros(6), sel)
            ok_o, state_o, iters_o, _ = oracle.vo_optimize(P, quads, np.zeros(6), sel)
            assert ok == ok_o and iters == iters_o
            assert rel(state, state_o) < 1e-6                                              # north_star: final poses within 1e-6


@pytest.mark.gpu
def test_gpu_ransac_matches_the_oracle_hypothesis_by_hypothesis(gpu_lib, oracle):
    """200 hypotheses (VisualOdometry.h:32) fitted and scored concurrently, 30 % outliers: per hypothesis the same
    accept / skip decision, the same inlier count and the same pose as the sequential oracle; the same winner; the final
    refinement over its consensus set lands on the oracle's motion."""
    quads, out, P, _ = _setup(1500, 0.3)
    rng = np.random.default_rng(17)
    triples = np.stack([rng.choice(len(quads), 3, replace=False) for _ in range(200)]).astype(np.int32)
    triples[7] = [4, 4, 9]                                   # a degenerate draw the caller could pass: skipped, not fatal
    o = oracle.vo_ransac(P, quads, np.zeros(6), np.where(triples == triples[7], triples, triples))
    with capi.Handle(capi.default_config(gpu_lib), lib=gpu_lib) as h:
        h.vo_set_matches(P, quads)
        g = h.vo_ransac(np.zeros(6), triples)
        assert np.array_equal(g["ok"], o["ok"]) and g["best"] == o["best"]
        assert np.abs(g["counts"] - o["counts"]).max() <= 1          # a match sitting on the threshold may flip with rounding
        good = o["ok"] == 1
        assert rel(g["states"][good], o["states"][good]) < 1e-6
        inl = h.vo_inliers()
        assert len(inl) == g["counts"][g["best"]] and out[inl].mean() < 0.02
        ok, state, iters = h.vo_refine(np.zeros(6))                 # over the resident consensus set
        ok_o, state_o, iters_o, _ = oracle.vo_optimize(P, quads, np.zeros(6), oracle.vo_inliers(P, quads, o["states"][o["best"]]))
        assert ok and ok_o and rel(state, state_o) < 1e-6 and abs(state[5] + 0.8) < 0.05
