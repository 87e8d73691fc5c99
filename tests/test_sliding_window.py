"""uba_window_advance — the per-frame sliding window (BASELINE config c2).

The reference re-runs initialiseObservations (BundleAdjuster.h:351-376) over the whole track container
(core/feature_types.h:121-191) for every frame.  Here the window slides on the device; the contract is that the handle
afterwards is INDISTINGUISHABLE from one that was given the equivalent window through uba_set_problem: same tables bit for
bit, same blocks, same trajectory — and the same answers as the oracle on that window.

CPU versions run on tests/emu (serial emulation: host logic + the thread-independent kernels), the -m gpu versions on
the product library."""
import numpy as np
import pytest

from uasl_motion_estimation_b200 import capi, synth

BLOCK_TOL = 1e-9
STATE_TOL = 1e-6
BLOCKS = ("residuals", "cost", "grad_cams", "B", "C", "S", "rhs")


def rel(a, b):
    a = np.asarray(a, float).reshape(-1); b = np.asarray(b, float).reshape(-1)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def fresh_handle(lib, win, **cfgkw):
    h = capi.Handle(capi.default_config(lib, **cfgkw), lib=lib)
    h.set_problem(4, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    return h


def slide_and_compare(lib, oracle, n_steps, optimise_each, n_frames=30, window=8, n_pts=4000, K=4, check_blocks=True):
    seq = synth.SlidingSequence(n_frames=n_frames, window=window, n_pts=n_pts, track_min=2, track_max=7, lib=lib if lib is not None else None)
    ids = seq.initial_ids()
    w0 = seq.window(0, ids)
    a = fresh_handle(lib, w0, sliding_window=1, fixed_iterations=K)
    if optimise_each:
        a.optimise(2)
    for first in range(n_steps):
        kw, ids_new = seq.advance(first, ids)
        old_n = a.n_pts
        id_map = a.window_advance(**kw)
        # pt_id_map: survivors keep their relative order, erased tracks say -1
        alive = seq.hi[ids] >= first + 1
        assert len(id_map) == old_n and np.array_equal(id_map >= 0, alive)
        assert np.array_equal(id_map[alive], np.arange(int(alive.sum())))
        ids = ids_new
        assert (a.n_cams, a.n_pts) == (window, len(ids))
        # the equivalent window, started from what the slid handle holds (previous solution + the newcomers)
        cams0, pts0 = a.cameras(), a.points()
        assert np.array_equal(pts0[int(alive.sum()):], kw["new_pts3"]) and np.array_equal(cams0[-1], kw["new_cams6"][0])
        wn = seq.window(first + 1, ids, cams=cams0, pts=pts0)
        assert a.n_obs == wn.n_obs
        b = fresh_handle(lib, wn, fixed_iterations=K)
        ta, tb = a.tables(2), b.tables(2)
        for k in ta:
            assert np.array_equal(ta[k], tb[k]), (k, first)
        r = oracle.tables(wn.n_cams, wn.n_pts, wn.cam_idx, wn.pt_idx, 2)
        for k in ta:
            assert np.array_equal(ta[k], r[k]), (k, first)
        if check_blocks:
            ga = a.linearize(2, 1e4, want=BLOCKS); gb = b.linearize(2, 1e4, want=BLOCKS)
            ro = oracle.linearize(wn, a.cfg, 2, 1e4)
            for k in BLOCKS:
                assert rel(ga[k], gb[k]) < 1e-12, (k, first)
                assert rel(ga[k], ro[k]) < BLOCK_TOL, (k, first)
        if optimise_each:
            rca, sa = a.optimise(2); rcb, sb = b.optimise(2)
            assert rca == rcb == 0
            assert rel(a.cameras(), b.cameras()) < 1e-9 and rel(a.points(), b.points()) < 1e-9
            o = oracle.optimise(wn, a.cfg, 2)
            assert rel(a.cameras(), o["cams"]) < STATE_TOL and rel(a.points(), o["pts"]) < STATE_TOL, first
            assert sa[0].iterations == K
        b.close()
    a.close()


def test_advance_matches_a_fresh_window_emu(emu_lib, oracle):
    slide_and_compare(emu_lib, oracle, n_steps=5, optimise_each=False)


def test_advance_after_optimise_matches_a_fresh_window_emu(emu_lib, oracle):
    slide_and_compare(emu_lib, oracle, n_steps=4, optimise_each=True, n_pts=2000)


def _errors(lib):
    seq = synth.SlidingSequence(n_frames=16, window=6, n_pts=800, track_min=2, track_max=5, lib=lib)
    ids = seq.initial_ids()
    w0 = seq.window(0, ids)
    kw, _ = seq.advance(0, ids)
    # not asked for at creation
    h = fresh_handle(lib, w0)
    with pytest.raises(capi.UbaError) as e:
        h.window_advance(**kw)
    assert e.value.code == capi.UBA_ERR_UNSUPPORTED
    h.close()
    h = fresh_handle(lib, w0, sliding_window=1)
    # an observation that skips a keyframe of its track
    bad = dict(kw); bad["n_drop"] = 0
    bad["new_cams6"] = np.concatenate([kw["new_cams6"], kw["new_cams6"]])
    bad["cam_idx"] = kw["cam_idx"] + 2
    with pytest.raises(capi.UbaError) as e:
        h.window_advance(**bad)
    assert e.value.code == capi.UBA_ERR_UNSUPPORTED
    # out-of-range indices
    bad = dict(kw); bad["pt_idx"] = kw["pt_idx"] + 10 ** 6
    with pytest.raises(capi.UbaError) as e:
        h.window_advance(**bad)
    assert e.value.code == capi.UBA_ERR_INVALID_ARGUMENT
    # the failed calls left the window as it was
    assert (h.n_cams, h.n_pts, h.n_obs) == (w0.n_cams, w0.n_pts, w0.n_obs)
    t = h.tables(2)
    assert np.array_equal(t["pt_obs_off"], np.r_[0, np.cumsum(np.bincount(w0.pt_idx, minlength=w0.n_pts))])
    h.window_advance(**kw)
    h.close()
    # a window whose tracks have gaps cannot slide
    keep = np.ones(w0.n_obs, bool)
    j = int(np.flatnonzero(np.bincount(w0.pt_idx) >= 3)[0]); o = np.flatnonzero(w0.pt_idx == j)
    keep[o[1]] = False
    h = capi.Handle(capi.default_config(lib, sliding_window=1), lib=lib)
    h.set_problem(4, w0.cams_init, w0.pts_init, w0.feats[keep], w0.cam_idx[keep], w0.pt_idx[keep], None, w0.calib)
    with pytest.raises(capi.UbaError) as e:
        h.window_advance(**kw)
    assert e.value.code == capi.UBA_ERR_UNSUPPORTED
    h.close()


def test_advance_errors_emu(emu_lib):
    _errors(emu_lib)


def test_advance_with_caller_supplied_iterate_emu(emu_lib, oracle):
    seq = synth.SlidingSequence(n_frames=16, window=6, n_pts=1500, track_min=2, track_max=5, lib=emu_lib)
    ids = seq.initial_ids()
    a = fresh_handle(emu_lib, seq.window(0, ids), sliding_window=1, fixed_iterations=3)
    a.optimise(2)
    kw, ids = seq.advance(0, ids)
    wn = seq.window(1, ids)                      # the generator's initial values for every camera / point of the new window
    a.window_advance(**kw, cams6_all=wn.cams_init, pts3_all=wn.pts_init)
    assert np.array_equal(a.cameras(), wn.cams_init) and np.array_equal(a.points(), wn.pts_init)
    a.optimise(2)
    o = oracle.optimise(wn, a.cfg, 2)
    assert rel(a.cameras(), o["cams"]) < STATE_TOL and rel(a.points(), o["pts"]) < STATE_TOL
    a.close()


# ---- the same on the GPU -------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_advance_matches_a_fresh_window_gpu(gpu_lib, oracle):
    slide_and_compare(gpu_lib, oracle, n_steps=6, optimise_each=True, n_frames=40, window=20, n_pts=20000, K=5)


@pytest.mark.gpu
def test_advance_errors_gpu(gpu_lib):
    _errors(gpu_lib)


@pytest.mark.gpu
def test_c2_sequence_every_20th_call_matches_the_oracle(gpu_lib, oracle):
    """BASELINE config c2 as it is defined: 200 keyframes, 20-keyframe window, one BA call per keyframe (181 calls), each
    warm-started from the previous call's solution.  Every 20th call is replayed by the oracle from the same start."""
    seq = synth.SlidingSequence()
    ids = seq.initial_ids()
    K = 4
    a = fresh_handle(gpu_lib, seq.window(0, ids), sliding_window=1, fixed_iterations=K)
    a.optimise(2)
    for first in range(seq.n_calls - 1):
        kw, ids = seq.advance(first, ids)
        a.window_advance(**kw)
        check = (first + 1) % 20 == 0 or first + 2 == seq.n_calls
        if check:
            cams0, pts0 = a.cameras(), a.points()
        rc, s = a.optimise(2)
        assert rc == 0 and s[0].usable
        if check:
            wn = seq.window(first + 1, ids, cams=cams0, pts=pts0)
            assert a.n_obs == wn.n_obs
            o = oracle.optimise(wn, a.cfg, 2)
            assert rel(a.cameras(), o["cams"]) < STATE_TOL and rel(a.points(), o["pts"]) < STATE_TOL, first
            assert abs(s[0].final_cost - o["summary"]["final_cost"]) <= 1e-9 * o["summary"]["final_cost"], first
    a.close()
