"""Pins the oracle (and the host-only boundary maths of libuba) to the REFERENCE'S OWN SOURCES.

Two sources of truth, both produced by code the reference ships (compiled in the build container by `make -C oracle
ref`, see oracle/ref_shim.cpp and oracle/refstub/):
  * oracle/_ref/libuba_ref.so, when present (it travels to the GPU box prebuilt), called live;
  * tests/golden/ref_golden.json, its committed outputs (exact hex floats), always.
What is pinned: residual rows and autodiff Jacobians of the three functors (BundleAdjuster.h:78-94,:113-130,:153-171),
log/exp map (rotation_utils.h:190-204), the observation table initialiseObservations builds (:351-376), parameter packing
(:297-310), and the whole optimise() (:431-476: bounds, fixed cameras, options, Status) driven through the reference class
with Ceres replaced by oracle/refstub's independent dense restatement."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import ref_binding as rb
from uasl_motion_estimation_b200 import capi, synth

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "ref_golden.json").read_text())
unhex = lambda a: np.array([float.fromhex(x) for x in a])
rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
needs_ref = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (no reference tree here)")


def _calib():
    k = capi.Calib()
    for n, v in GOLD["calib"].items():
        setattr(k, n, v)
    return k


def test_oracle_functor_rows_and_jacobians_equal_the_reference_bit_for_bit(oracle):
    """64 cases (stereo, mono left, mono right; rotations on both sides of the theta^2 = epsilon switch): the oracle's
    dual-number evaluation of its restated functors returns EXACTLY the doubles the reference's functor templates return
    through AutoDiffCostFunction."""
    k = _calib()
    for c in GOLD["functor"]:
        r, Jc, Jp = oracle.residual(c["M"], k, c["cam"], c["X"], c["obs"], c["cam_id"])
        assert np.array_equal(r, unhex(c["r"])), c
        assert np.array_equal(Jc.reshape(-1), unhex(c["Jc"])) and np.array_equal(Jp.reshape(-1), unhex(c["Jp"])), c


@needs_ref
def test_fixture_is_what_the_compiled_reference_returns_now():
    k = _calib()
    for c in GOLD["functor"][::5]:
        r, Jc, Jp = rb.residual(c["M"], k, c["cam"], c["X"], c["obs"], c["cam_id"])
        assert np.array_equal(r, unhex(c["r"])) and np.array_equal(Jc.reshape(-1), unhex(c["Jc"]))


def test_log_exp_map_of_libuba_and_oracle_equal_the_reference(oracle):
    """uba_log_map_quat / uba_exp_map_quat (libuba_host.so, what the adapter header calls) and the oracle's copies against
    log_map_Quat / exp_map_Quat of rotation_utils.h:190-204 (the Quat constructor normalises first, :120)."""
    host = capi.host_lib()
    for c in GOLD["quat"]:
        q = unhex(c["q"]); qn = q / np.linalg.norm(q)
        want = unhex(c["log"])
        for fn in (host.uba_log_map_quat, oracle.lib().uba_ref_log_map_quat):
            r = np.zeros(3); fn(capi.dptr(qn), capi.dptr(r))
            assert np.allclose(r, want, rtol=1e-13, atol=1e-15), (q, r, want)
        for fn in (host.uba_exp_map_quat, oracle.lib().uba_ref_exp_map_quat):
            qq = np.zeros(4); fn(capi.dptr(np.ascontiguousarray(want)), capi.dptr(qq))
            assert np.allclose(qq, unhex(c["exp_of_log"]), rtol=1e-13, atol=1e-15)


def test_log_map_clamps_where_the_reference_returns_nan():
    """log_map_Quat calls acos(w) unclamped (rotation_utils.h:203): w a hair above 1 is NaN there.  libuba clamps — a
    deliberate, documented difference (include/uba.h)."""
    host = capi.host_lib()
    q = np.array([1.0 + 4e-16, 0.0, 0.0, 0.0]); r = np.zeros(3)
    host.uba_log_map_quat(capi.dptr(q), capi.dptr(r))
    assert np.array_equal(r, np.zeros(3))
    assert np.isnan(np.arccos(q[0]))


@pytest.mark.parametrize("idx", range(len(GOLD["ba"])))
def test_reference_class_end_to_end_against_the_oracle(oracle, idx):
    """The reference's BundleAdjuster<M>, fed WBA points and CamPose_qd poses with frame IDs starting at 100, builds the
    observation table the synthetic generator / libuba / the oracle use (bit-exact), packs the poses to the same 6-vectors,
    and its optimise(fixedFrames) — default options of :463-467 — ends where the oracle's ends."""
    g = GOLD["ba"][idx]
    win = synth.config_window(g["config"], scale=g["scale"], M=g["M"])
    assert g["n_obs"] == win.n_obs
    assert np.array_equal(g["cam_idx"], win.cam_idx) and np.array_equal(g["pt_idx"], win.pt_idx) and np.array_equal(g["cam_id"], win.cam_id)
    t = oracle.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, g["fixed_frames"])
    assert np.array_equal(t["obs_order"], np.arange(win.n_obs))          # the reference's order IS the canonical one
    assert rel(unhex(g["cams_init"]).reshape(-1, 6), win.cams_init) < 1e-12     # log(exp(r)) round trip of the packing
    cfg = capi.default_config(max_solver_time_s=0.0)                       # the 1 s cap never binds at this size
    o = oracle.optimise(win, cfg, g["fixed_frames"])
    assert g["status"] == 2 and o["summary"]["usable"] == 1                 # Status::SUCCESSFUL
    assert rel(o["cams"], unhex(g["cams"]).reshape(-1, 6)) < 1e-9 and rel(o["pts"], unhex(g["pts"]).reshape(-1, 3)) < 1e-9
    qid = unhex(g["quat_id"]).reshape(-1, 5)
    assert np.array_equal(qid[:, 4], np.arange(win.n_cams))               # getCameraPoses renumbers IDs from 0 (:233-235)
    host = capi.host_lib()
    for c in range(win.n_cams):
        q = np.zeros(4); host.uba_exp_map_quat(capi.dptr(np.ascontiguousarray(o["cams"][c, 3:])), capi.dptr(q))
        assert np.allclose(q, qid[c, :4], atol=1e-9)


@needs_ref
def test_reference_class_drops_frames_before_the_window_and_fails_infeasible_starts(oracle):
    win = synth.config_window("c2", scale=0.003)
    # (a) tracks reach back before the first pose of the window: observations with frame_idx - first_frame < 0 are dropped
    #     (:370) while pt_idx keeps counting every track (:373)
    sub = synth.Window(win.M, win.cams_gt[3:], win.cams_init[3:], win.pts_gt, win.pts_init, win.feats, win.cam_idx, win.pt_idx,
                       win.cam_id, 2, win.calib)
    o = rb.ba_run(sub, first_frame=103 - 3 + 0, optimise=False)    # poses get IDs 100.., tracks are given in frames 100 + cam_idx
    # ba_run numbers frames from `first_frame` for BOTH; shift the poses instead: IDs 103.. -> first_frame = 103
    keep = win.cam_idx >= 3
    o = _run_shifted(sub, win, 3)
    assert np.array_equal(o["cam_idx"], win.cam_idx[keep] - 3) and np.array_equal(o["pt_idx"], win.pt_idx[keep])
    assert np.array_equal(o["feats"], win.feats[keep])
    # (b) a point outside its box: Ceres refuses the start, the class reports FAILED (3); the oracle says infeasible
    bad = synth.config_window("c1", scale=0.02)
    bad.pts_init[5, 2] = 1e6
    r = rb.ba_run(bad)
    assert r["status"] == 3
    oo = oracle.optimise(bad, capi.default_config(max_solver_time_s=0.0), 2)
    assert oo["rc"] == capi.UBA_ERR_INFEASIBLE and oo["summary"]["usable"] == 0


def _run_shifted(sub, win, shift):
    """sub has the poses of cameras [shift, n); the tracks still carry the frames of the full window."""
    M, nc, npt = win.M, sub.n_cams, win.n_pts
    poses = np.zeros((nc, 7))
    for c in range(nc):
        poses[c, :4] = rb.exp_map(sub.cams_init[c, 3:]); poses[c, 4:] = sub.cams_init[c, :3]
    cam_ids = np.arange(100 + shift, 100 + shift + nc, dtype=np.int32)
    counts = np.bincount(win.pt_idx, minlength=npt)
    sel = counts > 0
    track_off = np.concatenate([[0], np.cumsum(counts[sel])]).astype(np.int64)
    frame_idx = (win.cam_idx + 100).astype(np.int32)
    pts4 = np.concatenate([win.pts_init[sel], np.ones((int(sel.sum()), 1))], axis=1)
    assert sel.all()      # this window has no empty tracks, so pt_idx needs no remapping
    no = C.c_int64(0); mx = win.n_obs
    ci = np.zeros(mx, np.int32); pi = np.zeros(mx, np.int32); cid = np.zeros(mx, np.int32); ft = np.zeros((mx, M)); st = C.c_int32(0)
    rc = rb.lib().uba_refsrc_ba_run(M, C.byref(win.calib), 0, nc, capi.dptr(poses), capi.i32ptr(cam_ids), int(sel.sum()), capi.dptr(pts4), None,
                                    capi.i64ptr(track_off), capi.i32ptr(frame_idx), capi.dptr(win.feats), 2, 0, mx, C.byref(no),
                                    capi.i32ptr(ci), capi.i32ptr(pi), capi.i32ptr(cid), capi.dptr(ft), None, None, None, None, C.byref(st), None)
    assert rc == 0 and st.value == 1     # Status::INITIALISED
    n = no.value
    return dict(cam_idx=ci[:n], pt_idx=pi[:n], feats=ft[:n])
