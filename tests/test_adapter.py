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
