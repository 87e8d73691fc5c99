"""The drop-in me::optimisation::BundleAdjuster<M> header (uasl_motion_estimation_b200/include) compiled
against a minimal OpenCV/core stand-in (tests/cvstub) and driven like an application of the reference."""
import ctypes as C
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from uasl_motion_estimation_b200 import capi, synth

ROOT = Path(__file__).resolve().parent.parent
DEMO_DIR = ROOT / "tests" / "adapter"


def build_demo():
    lib_dir = ROOT / "uasl_motion_estimation_b200" / "lib"
    cmd = ["g++", "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'tests' / 'cvstub'}",
           f"-I{ROOT / 'uasl_motion_estimation_b200' / 'include'}", f"-I{ROOT / 'include'}", str(DEMO_DIR / "adapter_demo.cpp"),
           "-o", str(DEMO_DIR / "adapter_demo"), f"-L{lib_dir}", "-luba", f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(cmd, check=True, capture_output=True)
    return DEMO_DIR / "adapter_demo"


def test_adapter_compiles_as_cxx11_without_ceres():
    exe = build_demo()
    assert subprocess.run([str(exe), "--compile-only"]).returncode == 0
    text = (ROOT / "uasl_motion_estimation_b200" / "include" / "MotionEstimation" / "optimisation" / "BundleAdjuster.h").read_text()
    assert "#include <ceres" not in text and "google::" not in text


@pytest.mark.gpu
@pytest.mark.parametrize("M", [4, 2])
def test_adapter_matches_the_c_abi_path(gpu_lib, tmp_path, M):
    exe = build_demo()
    n_cams, n_pts, seed, fixed = 8, 300, 77, 2
    out = tmp_path / "ba.bin"
    r = subprocess.run([str(exe), str(M), str(n_cams), str(n_pts), str(seed), str(fixed), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "[Bundle Adjuster] optimising (8 cam poses and 300 pts with" in r.stdout
    assert "system should be initiliased" in r.stderr  # second optimise() refused, single use
    raw = out.read_bytes()
    status, nc, npt, no = struct.unpack("4i", raw[:16])
    poses = np.frombuffer(raw, np.float64, nc * 8, 16).reshape(nc, 8)
    pts = np.frombuffer(raw, np.float64, npt * 3, 16 + nc * 64).reshape(npt, 3)
    assert status == 2 and nc == n_cams and npt == n_pts  # Status::SUCCESSFUL
    np.testing.assert_array_equal(poses[:, 7], np.arange(nc))  # IDs renumbered from 0 (BundleAdjuster.h:233-235)
    win = synth.generate(n_cams, n_pts, 3, n_cams, seed=seed, M=M, fixed_frames=fixed, lib=gpu_lib)
    assert no == win.n_obs
    h = capi.Handle(capi.default_config(gpu_lib, compute_covariance=int(M == 4)), lib=gpu_lib)
    h.set_problem(M, win.cams_init, win.pts_init, win.feats, win.cam_idx, win.pt_idx, win.cam_id, win.calib)
    rc, sums = h.optimise(fixed)
    assert rc == 0
    cams = h.cameras()
    q = np.zeros((nc, 4))
    for i in range(nc):
        gpu_lib.uba_exp_map_quat(capi.dptr(np.ascontiguousarray(cams[i, 3:])), capi.dptr(q[i]))
    np.testing.assert_allclose(poses[:, :4], q, rtol=0, atol=1e-8)
    np.testing.assert_allclose(poses[:, 4:7], cams[:, :3], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(pts, h.points(), rtol=1e-7, atol=1e-7)
    off = 16 + nc * 64 + npt * 24
    (ncov,) = struct.unpack("i", raw[off:off + 4])
    if M == 4:  # CalibrationParameters::compute_cov -> getPosesCovariance()
        assert ncov == nc
        cov = np.frombuffer(raw, np.float64, nc * 36, off + 4).reshape(nc, 6, 6)
        ref = h.pose_covariances()
        np.testing.assert_allclose(cov, ref, rtol=1e-5, atol=1e-12)
        assert not cov[:fixed].any() and cov[fixed:].any()
    else:
        assert ncov == 0


# ---------------------------------------------------------------------------------------------------------------------
# me::StereoVisualOdometry drop-in (uasl_motion_estimation_b200/include/MotionEstimation/vo/StereoVisualOdometry.h)
def build_vo_demo():
    lib_dir = ROOT / "uasl_motion_estimation_b200" / "lib"
    inc = ROOT / "uasl_motion_estimation_b200" / "include"
    cmd = ["g++", "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'tests' / 'cvstub'}",
           f"-I{inc / 'MotionEstimation'}", f"-I{inc}", f"-I{ROOT / 'include'}", str(DEMO_DIR / "vo_demo.cpp"),
           "-o", str(DEMO_DIR / "vo_demo"), f"-L{lib_dir}", "-luba", f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(cmd, check=True, capture_output=True)
    return DEMO_DIR / "vo_demo"


def test_vo_dropin_compiles_as_cxx11_with_the_reference_interface():
    exe = build_vo_demo()
    assert subprocess.run([str(exe), "--compile-only"]).returncode == 0
    text = (ROOT / "uasl_motion_estimation_b200" / "include" / "MotionEstimation" / "vo" / "StereoVisualOdometry.h").read_text()
    for member in ("bool process(const std::vector<StereoOdoMatchesf>& matches, cv::Mat init", "virtual cv::Mat getMotion()",
                   "getPts3D()", "getInliers_idx()", "getPredictions()", "getParams()"):
        assert member in text


def test_vo_pose_matrix_and_predictions_follow_the_oracle(oracle):
    """The host helpers behind getMotion() / getPredictions(): [R(euler)^T | t] and reproject(), against the oracle's
    residuals (observed - predicted)."""
    host = capi.load_host()
    quads, _ = synth.vo_quads(50, 0.0)
    P = synth.vo_params(capi.default_calib())
    state = np.array([0.01, -0.02, 0.015, 0.1, -0.05, -0.8])
    T = np.zeros(16); host.uba_vo_pose_matrix(capi.dptr(state), capi.dptr(T)); T = T.reshape(4, 4)
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-14) and np.array_equal(T[:3, 3], state[3:]) and np.array_equal(T[3], [0, 0, 0, 1])
    pts = oracle.vo_project3d(P, quads)
    pred = np.zeros((len(quads), 4))
    host.uba_vo_predict(C.byref(P), capi.dptr(state), len(quads), capi.dptr(np.ascontiguousarray(pts)), capi.dptr(pred))
    sel = np.arange(len(quads), dtype=np.int32)
    res = oracle.vo_linearize(P, quads, state, sel)["res"]           # observed - predicted, [n][4]
    obs = np.asarray(quads, np.float64).reshape(-1, 8)[:, 4:8]
    assert np.abs((obs - pred) - res).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("ransac,method", [(1, 0), (1, 1), (0, 0)])
def test_vo_dropin_matches_the_c_abi_path(gpu_lib, oracle, tmp_path, ransac, method):
    """The class driven like the reference's (srand + process + getters) against the same steps through the C ABI with the
    triples rand() produces for that seed, and against the oracle."""
    exe = build_vo_demo()
    n, seed, n_ransac = 500, 1234, 50
    quads, out = synth.vo_quads(n, 0.25 if ransac else 0.0)
    calib = capi.default_calib()
    P = synth.vo_params(calib, method=method)
    init = np.zeros(6)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<5i", n, seed, ransac, n_ransac, method))
        f.write(struct.pack("<6d", P.fu1, P.fv1, P.cu1, P.cv1, P.baseline, P.inlier_threshold))
        f.write(struct.pack("<6d", *init))
        f.write(np.ascontiguousarray(quads, np.float32).tobytes())
    r = subprocess.run([str(exe), str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = fout.read_bytes()
    ok, n_in = struct.unpack_from("<2i", raw, 0)
    off = 8
    T = np.frombuffer(raw, np.float64, 16, off).reshape(4, 4); off += 128
    inl = np.frombuffer(raw, np.int32, n_in, off); off += 4 * n_in
    pts = np.frombuffer(raw, np.float64, 4 * n, off).reshape(n, 4); off += 32 * n
    pred = np.frombuffer(raw, np.float64, 4 * n_in, off).reshape(n_in, 4)
    assert f"[Motion Estimation] {n_in} inliers" in r.stdout

    # the same through the C ABI: rand() of the C library with the same seed gives the reference's triples
    libc = C.CDLL("libc.so.6")
    libc.srand(seed)
    triples = []
    for _ in range(n_ransac if ransac else 0):
        t = []
        while len(t) < 3:
            idx = libc.rand() % n
            if idx not in t:
                t.append(idx)
        triples.append(t)
    with capi.Handle(capi.default_config(gpu_lib), lib=gpu_lib) as h:
        h.vo_set_matches(P, quads)
        assert np.array_equal(h.vo_points(), pts)
        if ransac:
            h.vo_ransac(init, np.array(triples, np.int32))
            ref_inl = h.vo_inliers()
        else:
            ref_inl = np.arange(n, dtype=np.int32)
        assert np.array_equal(ref_inl, inl)
        conv, state, iters = h.vo_refine(init, ref_inl)
    assert bool(ok) == bool(conv) and ok == 1
    Tref = np.zeros(16); gpu_lib.uba_vo_pose_matrix(capi.dptr(state), capi.dptr(Tref))
    assert np.array_equal(Tref.reshape(4, 4), T)
    assert abs(T[2, 3] + 0.8) < 0.05                                   # the rig advances 0.8 m per keyframe
    if ransac:
        assert out[inl].mean() < 0.02
    # and the oracle agrees with the refined motion and the predictions
    ok_o, state_o, _, _ = oracle.vo_optimize(P, quads, init, inl.astype(np.int32))
    assert ok_o and np.abs(state_o - state).max() < 1e-8
    res = oracle.vo_linearize(P, quads, state, inl.astype(np.int32))["res"]
    obs = np.asarray(quads, np.float64).reshape(-1, 8)[inl][:, 4:8]
    assert np.abs((obs - pred) - res).max() < 1e-8
