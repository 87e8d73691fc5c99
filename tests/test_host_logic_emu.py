"""Host logic of libuba (index tables, launch sequencing, LM controller, error behaviour) exercised on
the CPU through tests/emu — a serial emulation of the thread-independent kernels.  The parity tests
proper are the -m gpu tests in test_gpu_parity.py; these make the same comparisons at small sizes so
host-side regressions are caught without a GPU."""
import numpy as np
import pytest

from uasl_motion_estimation_b200 import capi, synth


BLOCK_TOL = 1e-9   # north_star: residual vectors and normal-equation blocks within 1e-9 relative
STATE_TOL = 1e-6   # final poses and points within 1e-6 relative after a fixed number of LM iterations
BLOCKS = ("residuals", "weights", "cost", "grad_cams", "grad_pts", "B", "C", "W", "S", "rhs", "lm_diag_cams", "lm_diag_pts")


def rel(a, b):
    a = np.asarray(a, float).reshape(-1); b = np.asarray(b, float).reshape(-1)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def make(lib, name, scale, M=4, **cfgkw):
    win = synth.config_window(name, scale=scale, lib=lib, M=M)
    cfg = capi.default_config(lib, loss_kind=synth.CONFIGS[name]["loss"], **cfgkw)
    h = capi.Handle(cfg, lib=lib)
    h.set_problem(M, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    return win, cfg, h


@pytest.mark.parametrize("name,scale", [("c1", 0.02), ("c2", 0.004), ("c4", 0.001), ("c5", 0.002)])
def test_tables_bit_exact(emu_lib, oracle, name, scale):
    win, cfg, h = make(emu_lib, name, scale)
    for fixed in (0, 2, 3):
        t = h.tables(fixed); r = oracle.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, fixed)
        for k in t:
            assert np.array_equal(t[k], r[k]), (k, fixed)
    # the reference's own ordering (point-major, frame-ascending) is the identity
    assert np.array_equal(h.tables(2)["obs_order"], np.arange(win.n_obs))


def test_tables_with_shuffled_observations_and_empty_points(emu_lib, oracle):
    win = synth.config_window("c2", scale=0.004, lib=emu_lib)
    rng = np.random.default_rng(0)
    perm = rng.permutation(win.n_obs)
    keep = perm[win.pt_idx[perm] % 7 != 3]  # points 3, 10, 17, ... lose every observation
    ci, pi, f = win.cam_idx[keep], win.pt_idx[keep], win.feats[keep]
    h = capi.Handle(capi.default_config(emu_lib), lib=emu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, f, ci, pi, None, win.calib)
    t = h.tables(2); r = oracle.tables(win.n_cams, win.n_pts, ci, pi, 2)
    for k in t:
        assert np.array_equal(t[k], r[k]), k
    assert (np.diff(t["pt_obs_off"])[3::7] == 0).all()


def test_tables_of_a_large_window_with_empty_points(emu_lib, oracle):
    """Above 65536 points the offsets come from chunked parallel passes (backward fill over points without observations,
    exclusive scan of the counts): same tables as the oracle, with runs of empty points across the chunk borders."""
    win = synth.config_window("c4", scale=0.5, lib=emu_lib)
    assert win.n_pts > 65536
    empty = np.zeros(win.n_pts, bool)
    empty[::11] = True; empty[12000:13000] = True; empty[win.n_pts - 700:] = True; empty[:3] = True
    keep = ~empty[win.pt_idx]                                   # point-major order kept: the sorted-input path
    ci, pi, f = win.cam_idx[keep], win.pt_idx[keep], win.feats[keep]
    h = capi.Handle(capi.default_config(emu_lib), lib=emu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, f, ci, pi, None, win.calib)
    t = h.tables(2); r = oracle.tables(win.n_cams, win.n_pts, ci, pi, 2)
    for k in t:
        assert np.array_equal(t[k], r[k]), k
    assert (np.diff(t["pt_obs_off"])[empty] == 0).all() and (np.diff(t["pt_obs_off"])[~empty] > 0).all()


@pytest.mark.parametrize("variant", ["track_order", "camera_descending", "random_points"])
def test_ingest_paths_give_the_same_tables_and_blocks(emu_lib, oracle, variant):
    """The ingest has a fast path for input that is already canonical (tracks grouped by first/last keyframe, observations
    point-major and frame-ascending: nothing moves) next to the general one (parallel counting sort of the points, per-point
    observation sort).  Each must give the oracle's tables bit for bit and its blocks to 1e-9."""
    base = synth.config_window("c2", scale=0.006, lib=emu_lib)
    if variant == "track_order":
        win = synth.reorder(base)
    elif variant == "camera_descending":
        win = synth.reorder(base, camera_descending=True)
    else:
        win = synth.reorder(base, point_order=np.random.default_rng(3).permutation(base.n_pts))
    cfg = capi.default_config(emu_lib, fixed_iterations=3)
    h = capi.Handle(cfg, lib=emu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    t = h.tables(2); r = oracle.tables(win.n_cams, win.n_pts, win.cam_idx, win.pt_idx, 2)
    for k in t:
        assert np.array_equal(t[k], r[k]), k
    if variant == "track_order":
        assert np.array_equal(t["obs_order"], np.arange(win.n_obs)) and np.array_equal(t["pt_order"], np.arange(win.n_pts))
    g = h.linearize(2, 1e4); o = oracle.linearize(win, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], o[k]) < BLOCK_TOL, k
    rc, _ = h.optimise(2)
    oo = oracle.optimise(win, cfg, 2)
    assert rc == 0 and rel(h.cameras(), oo["cams"]) < STATE_TOL and rel(h.points(), oo["pts"]) < STATE_TOL


@pytest.mark.parametrize("name,scale,M", [("c1", 0.02, 4), ("c2", 0.004, 4), ("c4", 0.001, 4), ("c5", 0.002, 4), ("c1", 0.02, 2), ("c2", 0.004, 2)])
def test_linearization_blocks(emu_lib, oracle, name, scale, M):
    win, cfg, h = make(emu_lib, name, scale, M)
    g = h.linearize(2, 1e4); r = oracle.linearize(win, cfg, 2, 1e4)
    for k in ("residuals", "weights", "cost", "grad_cams", "grad_pts", "B", "C", "W", "S", "rhs", "lm_diag_cams", "lm_diag_pts"):
        assert rel(g[k], r[k]) < 1e-9, k


@pytest.mark.parametrize("name,scale,M", [("c1", 0.02, 4), ("c2", 0.004, 4), ("c4", 0.001, 4), ("c1", 0.02, 2)])
def test_fixed_iteration_trajectory(emu_lib, oracle, name, scale, M):
    win, cfg, h = make(emu_lib, name, scale, M, fixed_iterations=6)
    rc, sums = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    assert rc == 0 and o["rc"] == 0
    assert rel(h.cameras(), o["cams"]) < 1e-6 and rel(h.points(), o["pts"]) < 1e-6
    gi = h.iterations(0)
    assert len(gi) == len(o["iterations"]) == 7
    for a, b in zip(gi, o["iterations"]):
        assert a["accepted"] == b["accepted"]
        assert a["cost"] == pytest.approx(b["cost"], rel=1e-7)
        assert a["radius"] == pytest.approx(b["radius"], rel=1e-5)
        assert a["model_cost_change"] == pytest.approx(b["model_cost_change"], rel=1e-6, abs=1e-12)


def test_reference_termination_rules(emu_lib, oracle):
    """Default options (function_tolerance 1e-3, Ceres defaults): same termination, same iterate."""
    win, cfg, h = make(emu_lib, "c1", 0.02, max_solver_time_s=0.0)
    rc, sums = h.optimise(2)
    o = oracle.optimise(win, cfg, 2)
    s = sums[0].as_dict()
    assert s["termination"] == o["summary"]["termination"] == 1  # CONVERGENCE_FUNCTION
    assert s["iterations"] == o["summary"]["iterations"]
    assert rel(h.cameras(), o["cams"]) < 1e-6 and rel(h.points(), o["pts"]) < 1e-6


def test_batch_of_windows_matches_single_windows(emu_lib, oracle):
    wins = [synth.config_window("c1", window=i, scale=0.01 + 0.002 * i, lib=emu_lib) for i in range(4)]
    cfg = capi.default_config(emu_lib, max_solver_time_s=0.0)
    h = capi.Handle(cfg, lib=emu_lib)
    h.set_batch(**synth.concat_windows(wins))
    rc, sums = h.optimise(2)
    assert rc == 0
    cams = h.cameras(); pts = h.points()
    c0 = p0 = 0
    for w, win in enumerate(wins):
        o = oracle.optimise(win, cfg, 2)
        assert sums[w].iterations == o["summary"]["iterations"] and sums[w].termination == o["summary"]["termination"]
        assert rel(cams[c0:c0 + win.n_cams], o["cams"]) < 1e-6
        assert rel(pts[p0:p0 + win.n_pts], o["pts"]) < 1e-6
        c0 += win.n_cams; p0 += win.n_pts


def test_state_machine_and_errors(emu_lib):
    h = capi.Handle(capi.default_config(emu_lib), lib=emu_lib)
    with pytest.raises(capi.UbaError) as e:
        h.n_windows = 1; h.optimise(2)
    assert e.value.code == capi.UBA_ERR_STATE  # "system should be initialised" (BundleAdjuster.h:381-384)
    win = synth.config_window("c1", scale=0.01, lib=emu_lib)
    bad = win.cam_idx.copy(); bad[5] = win.n_cams  # the reference has no upper-bound check (OOB read)
    with pytest.raises(capi.UbaError) as e:
        h.set_problem(4, win.cams_init, win.pts_init, win.feats, bad, win.pt_idx, None, win.calib)
    assert e.value.code == capi.UBA_ERR_INVALID_ARGUMENT
    with pytest.raises(capi.UbaError):
        h.set_problem(3, win.cams_init, win.pts_init, win.feats[:, :3], win.cam_idx, win.pt_idx, None, win.calib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, None, win.calib)
    np.testing.assert_array_equal(h.points(), win.pts_init)  # getters before optimise return the inputs
    rc, _ = h.optimise(2)
    assert rc == 0
    with pytest.raises(capi.UbaError) as e:  # single use, like the reference (BundleAdjuster.h:7)
        h.optimise(2)
    assert e.value.code == capi.UBA_ERR_STATE


def test_infeasible_start_is_a_failure_and_restores_inputs(emu_lib):
    win = synth.config_window("c1", scale=0.01, lib=emu_lib)
    win.pts_init[3, 2] = 1e9
    h = capi.Handle(capi.default_config(emu_lib), lib=emu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, None, win.calib)
    rc, sums = h.optimise(2, check=False)
    assert rc == capi.UBA_ERR_INFEASIBLE and sums[0].usable == 0 and sums[0].termination == 5
    np.testing.assert_array_equal(h.points(), win.pts_init)
    np.testing.assert_array_equal(h.cameras(), win.cams_init)


def test_all_cameras_fixed_moves_only_points(emu_lib, oracle):
    win, cfg, h = make(emu_lib, "c1", 0.01, fixed_iterations=3)
    rc, _ = h.optimise(win.n_cams)
    o = oracle.optimise(win, cfg, win.n_cams)
    np.testing.assert_array_equal(h.cameras(), win.cams_init)
    assert rel(h.points(), o["pts"]) < 1e-6


def test_features_that_are_not_floats_take_the_double_upload(emu_lib, oracle):
    """The reference's features are float detections widened to double, and such rows travel to the GPU as float32 (exactly).
    Rows with any value that is not a float must take the double path and still match the oracle on the same doubles."""
    base = synth.config_window("c2", scale=0.01, lib=emu_lib)
    feats = base.feats + 1e-7 * np.sin(np.arange(base.feats.size)).reshape(base.feats.shape)   # no longer float32-representable
    assert (feats.astype(np.float32).astype(np.float64) != feats).any()
    win = synth.Window(4, base.cams_gt, base.cams_init, base.pts_gt, base.pts_init, np.ascontiguousarray(feats), base.cam_idx, base.pt_idx,
                       base.cam_id, 2, base.calib)
    cfg = capi.default_config(emu_lib, fixed_iterations=3)
    h = capi.Handle(cfg, lib=emu_lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    g = h.linearize(2, 1e4); o = oracle.linearize(win, cfg, 2, 1e4)
    for k in BLOCKS:
        assert rel(g[k], o[k]) < BLOCK_TOL, k
    # the residuals see the perturbation: they differ from the float-rounded window's by more than the parity tolerance
    hb = capi.Handle(cfg, lib=emu_lib)
    hb.set_problem(4, base.cams_init, base.pts_init, base.feats, base.cam_idx, base.pt_idx, base.cam_id, base.calib)
    assert rel(hb.linearize(2, 1e4)["residuals"], g["residuals"]) > 1e-9


def _tile_plan(lib, h, fixed):
    import ctypes as C
    fn = lib.uba_emu_tile_plan
    fn.restype = C.c_int
    parts = np.zeros((8192, 7), np.int32); mask = np.zeros(h.n_pts, np.uint32)
    n_parts = C.c_int32(0); n_gen = C.c_int32(0)
    rc = fn(h._h, fixed, parts.ctypes.data_as(C.c_void_p), parts.shape[0], mask.ctypes.data_as(C.c_void_p), C.byref(n_parts), C.byref(n_gen))
    assert rc == 0 and n_parts.value <= parts.shape[0]
    return parts[: n_parts.value], mask, n_gen.value


@pytest.mark.parametrize("name,scale", [("c4", 1.0), ("c4", 0.05), ("c5", 1.0), ("c1", 1.0), ("c2", 0.2)])
def test_tile_plan_covers_every_track_once(emu_lib, name, scale):
    """The lineariser's plan (build_tile_plan): parts are disjoint runs of internal point slots that cover every observed point
    once; the slot mask of a point has one bit per observation, inside the part's camera list; parts of one kernel are handed
    out longest first; a slot-kernel part has at most 10 local cameras; and the model's choice on the BASELINE shapes is the
    one DESIGN.md states (c4: every part within one chunk of the others, about two full rounds of the 296 resident CTAs)."""
    win, cfg, h = make(emu_lib, name, scale)
    fixed = 2
    parts, mask, n_gen = _tile_plan(emu_lib, h, fixed)
    t = h.tables(fixed)
    track = np.diff(t["pt_obs_off"]).astype(np.int64)[t["pt_order"]]     # caller order -> internal slots
    assert n_gen == 0 and len(parts) > 0
    cover = np.zeros(h.n_pts, np.int32)
    for w, b, e, nl, nfx, slot, c0 in parts:
        assert 0 <= b < e <= h.n_pts and 1 <= nl <= 32 and 0 <= nfx <= nl
        cover[b:e] += 1
        m = mask[b:e]
        assert (m >> np.uint32(nl) == 0).all() if nl < 32 else True      # no bit beyond the part's camera list
        if slot:
            assert nl <= 10
    observed = track > 0
    assert (cover[observed] == 1).all() and (cover <= 1).all()
    pop = np.array([bin(int(x)).count("1") for x in mask[observed]])
    assert (pop == track[observed]).all()
    # longest first inside a kernel class (slot parts: up to 5 cameras, up to 10; the others keep their own classes)
    slot_parts = parts[parts[:, 5] == 1]
    for cls in (0, 1):
        sel = slot_parts[(slot_parts[:, 3] > 5) == (cls == 1)]
        length = sel[:, 2] - sel[:, 1]
        assert (np.diff(length) <= 0).all()
    if name == "c4" and scale == 1.0:
        chunks = (slot_parts[:, 2] - slot_parts[:, 1] + 31) // 32
        assert len(slot_parts) == len(parts) and 560 <= len(parts) <= 592 and chunks.max() - chunks.min() <= 3
