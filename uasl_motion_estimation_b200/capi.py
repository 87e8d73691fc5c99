"""ctypes binding of libuba's C ABI (include/uba.h).

The library is the product; this module is plumbing for tests, bench.py and the Python
mirror of the reference interface.  There is no Python or CPU implementation behind it:
if ``lib/libuba.so`` is missing the import of :func:`load` fails loudly, and
``uba_create`` fails with ``UBA_ERR_CUDA`` on a box without a B200.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
DEFAULT_LIB = PKG_DIR / "lib" / "libuba.so"
HOST_LIB = PKG_DIR / "lib" / "libuba_host.so"   # host-only part of the ABI (generator, defaults, boundary maths): no CUDA

UBA_OK = 0
UBA_ERR_INVALID_ARGUMENT = -1
UBA_ERR_STATE = -2
UBA_ERR_CUDA = -3
UBA_ERR_INFEASIBLE = -4
UBA_ERR_NUMERICAL = -5
UBA_ERR_NCCL = -6
UBA_ERR_UNSUPPORTED = -7

LOSS_TRIVIAL, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2
TERM_NAMES = {0: "RUNNING", 1: "CONVERGENCE_FUNCTION", 2: "CONVERGENCE_GRADIENT", 3: "CONVERGENCE_PARAMETER",
              4: "NO_CONVERGENCE", 5: "FAILURE", 6: "CONVERGENCE_RADIUS"}
NCCL_UNIQUE_ID_BYTES = 128

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Calib(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fx0", "fy0", "cx0", "cy0", "fx1", "fy1", "cx1", "cy1", "feat_var", "baseline")]


class Config(C.Structure):
    _fields_ = [("loss_kind", C.c_int32), ("loss_scale", C.c_double), ("max_iterations", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_consecutive_invalid_steps", C.c_int32), ("max_solver_time_s", C.c_double),
                ("fixed_iterations", C.c_int32), ("jacobi_scaling", C.c_int32), ("use_bounds", C.c_int32),
                ("device", C.c_int32), ("linearizer", C.c_int32), ("compute_covariance", C.c_int32), ("solver", C.c_int32),
                ("sliding_window", C.c_int32)]


class Summary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("usable", C.c_int32), ("iterations", C.c_int32), ("successful_steps", C.c_int32),
                ("unsuccessful_steps", C.c_int32), ("invalid_steps", C.c_int32), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("final_radius", C.c_double), ("final_gradient_max_norm", C.c_double)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["termination_name"] = TERM_NAMES.get(self.termination, "?")
        return d


class Iteration(C.Structure):
    _fields_ = [("cost", C.c_double), ("candidate_cost", C.c_double), ("model_cost_change", C.c_double),
                ("relative_decrease", C.c_double), ("radius", C.c_double), ("step_norm", C.c_double),
                ("gradient_max_norm", C.c_double), ("accepted", C.c_int32), ("pad_", C.c_int32)]


class LinearizationOut(C.Structure):
    _fields_ = [(n, c_double_p) for n in ("residuals", "weights", "cost", "grad_cams", "grad_pts", "B", "C", "W", "S", "rhs",
                                          "lm_diag_cams", "lm_diag_pts")]


class Timing(C.Structure):
    _fields_ = [("linearize_ms", C.c_double), ("solve_ms", C.c_double), ("backsub_ms", C.c_double), ("update_ms", C.c_double),
                ("comm_ms", C.c_double), ("total_ms", C.c_double), ("linearize_launches", C.c_int64), ("kernel_launches", C.c_int64)]


class VoParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fu1", "fv1", "cu1", "cv1", "fu2", "fv2", "cu2", "cv2", "baseline")] + \
               [("method", C.c_int32), ("max_iter", C.c_int32)] + [(n, C.c_double) for n in ("e1", "e2", "e3", "e4", "inlier_threshold")]


class SynthSpec(C.Structure):
    _fields_ = [("M", C.c_int32), ("n_cams", C.c_int32), ("n_pts", C.c_int32), ("track_min", C.c_int32), ("track_max", C.c_int32),
                ("full_tracks", C.c_int32), ("outlier_fraction", C.c_double), ("pixel_sigma", C.c_double),
                ("pose_t_sigma", C.c_double), ("pose_r_sigma", C.c_double), ("point_rel_sigma", C.c_double),
                ("fixed_frames", C.c_int32), ("seed", C.c_uint64)]


# every entry point include/uba.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("uba_config_default", None, [C.POINTER(Config)]),
    ("uba_create", C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    ("uba_destroy", None, [C.c_void_p]),
    ("uba_last_error", C.c_char_p, [C.c_void_p]),
    ("uba_version", C.c_int, []),
    ("uba_set_problem", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                   c_int32_p, c_int32_p, c_int32_p, C.POINTER(Calib)]),
    ("uba_set_batch", C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int32_p, c_int32_p, c_int64_p, c_double_p, c_double_p, c_double_p,
                                 c_int32_p, c_int32_p, c_int32_p, C.POINTER(Calib)]),
    ("uba_window_advance", C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, c_int32_p, C.c_int, c_double_p, c_int32_p,
                                      c_int32_p, c_double_p, c_double_p, c_int32_p]),
    ("uba_linearize", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.POINTER(LinearizationOut)]),
    ("uba_optimise", C.c_int, [C.c_void_p, C.c_int, C.POINTER(Summary)]),
    ("uba_get_cameras", C.c_int, [C.c_void_p, c_double_p]),
    ("uba_get_points", C.c_int, [C.c_void_p, c_double_p]),
    ("uba_get_pose_covariances", C.c_int, [C.c_void_p, c_double_p]),
    ("uba_get_iterations", C.c_int, [C.c_void_p, C.c_int, C.POINTER(Iteration), C.c_int, C.POINTER(C.c_int)]),
    ("uba_get_sizes", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    ("uba_get_tables", C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_int64_p, c_int32_p, c_int32_p]),
    ("uba_comm_unique_id", C.c_int, [C.c_void_p, C.c_char_p]),
    ("uba_comm_init", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    ("uba_set_profiling", C.c_int, [C.c_void_p, C.c_int]),
    ("uba_get_timing", C.c_int, [C.c_void_p, C.POINTER(Timing), C.c_int]),
    ("uba_time_linearize", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p]),
    ("uba_time_iteration", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    ("uba_probe_fp64_tflops", C.c_int, [C.c_void_p, c_double_p]),
    ("uba_synth_default_calib", None, [C.POINTER(Calib)]),
    ("uba_synth_generate", C.c_int64, [C.POINTER(SynthSpec), C.POINTER(Calib), C.c_int64, c_double_p, c_double_p, c_double_p,
                                        c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p]),
    ("uba_vo_params_default", None, [C.POINTER(VoParams)]),
    ("uba_vo_set_matches", C.c_int, [C.c_void_p, C.POINTER(VoParams), C.c_int, C.POINTER(C.c_float)]),
    ("uba_vo_get_points", C.c_int, [C.c_void_p, c_double_p]),
    ("uba_vo_linearize", C.c_int, [C.c_void_p, c_double_p, C.c_int, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("uba_vo_ransac", C.c_int, [C.c_void_p, c_double_p, C.c_int, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_double_p]),
    ("uba_vo_get_inliers", C.c_int, [C.c_void_p, c_int32_p, c_int32_p]),
    ("uba_vo_refine", C.c_int, [C.c_void_p, c_double_p, C.c_int, c_int32_p, c_double_p, c_int32_p, c_int32_p]),
    ("uba_vo_pose_matrix", None, [c_double_p, c_double_p]),
    ("uba_vo_predict", None, [C.POINTER(VoParams), c_double_p, C.c_int, c_double_p, c_double_p]),
    ("uba_shard_points", C.c_int, [C.c_int, C.c_int, C.c_int64, c_int32_p, c_int32_p, C.c_int, c_int32_p, c_int64_p, c_int32_p]),
    ("uba_shard_extract", C.c_int64, [C.c_int, C.c_int, C.c_int64, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                       C.c_int, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    ("uba_log_map_quat", None, [c_double_p, c_double_p]),
    ("uba_exp_map_quat", None, [c_double_p, c_double_p]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]
# the subset libuba_host.so carries as well (same sources, compiled without CUDA)
HOST_SYMBOLS = ["uba_config_default", "uba_synth_default_calib", "uba_synth_generate", "uba_log_map_quat", "uba_exp_map_quat",
                "uba_shard_points", "uba_shard_extract", "uba_vo_params_default", "uba_vo_pose_matrix", "uba_vo_predict"]


class UbaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libuba error {code}: {message}")
        self.code = code


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def i32ptr(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def i64ptr(a):
    return None if a is None else a.ctypes.data_as(c_int64_p)


def load(path: str | os.PathLike | None = None, allow_missing: tuple = ()) -> C.CDLL:
    """Load libuba.  Fails loudly when the CUDA library has not been built.  `allow_missing` names entry points a TEST build
    (tests/emu, which cannot emulate the shared-memory kernels) may lack; the product library must export everything."""
    p = Path(path) if path is not None else DEFAULT_LIB
    if not p.exists():
        raise ImportError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(nvcc, sm_100a).  uasl_motion_estimation_b200 has no CPU implementation.")
    lib = C.CDLL(str(p))
    for name, res, args in _SIGNATURES:
        if name in allow_missing and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def load_host(path: str | os.PathLike | None = None) -> C.CDLL:
    """Load libuba_host.so: generator, defaults and boundary maths only.  Nothing in it computes bundle adjustment."""
    p = Path(path) if path is not None else HOST_LIB
    if not p.exists():
        raise ImportError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(str(p))
    for name, res, args in _SIGNATURES:
        if name in HOST_SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    return lib


_default_lib = None
_host_lib = None


def default_lib() -> C.CDLL:
    global _default_lib
    if _default_lib is None:
        _default_lib = load()
    return _default_lib


def host_lib() -> C.CDLL:
    global _host_lib
    if _host_lib is None:
        _host_lib = load_host()
    return _host_lib


def default_config(lib=None, **overrides) -> Config:
    lib = lib or host_lib()
    cfg = Config()
    lib.uba_config_default(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"uba_config has no field {k}")
        setattr(cfg, k, v)
    return cfg


def default_calib(lib=None) -> Calib:
    lib = lib or host_lib()
    k = Calib()
    lib.uba_synth_default_calib(C.byref(k))
    return k


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Handle:
    """Thin RAII wrapper over uba_handle."""

    def __init__(self, cfg: Config | None = None, lib=None, **overrides):
        self.lib = lib or default_lib()
        self.cfg = cfg if cfg is not None else default_config(self.lib, **overrides)
        self._h = C.c_void_p()
        rc = self.lib.uba_create(C.byref(self.cfg), C.byref(self._h))
        if rc != UBA_OK:
            raise UbaError(rc, (self.lib.uba_last_error(None) or b"").decode())
        self.M = 0
        self.n_windows = self.n_cams = self.n_pts = self.n_obs = 0
        self._keep = []

    def close(self):
        if self._h:
            self.lib.uba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, ok=(UBA_OK,)):
        if rc not in ok:
            raise UbaError(rc, (self.lib.uba_last_error(self._h) or b"").decode())
        return rc

    def last_error(self) -> str:
        return (self.lib.uba_last_error(self._h) or b"").decode()

    # ---- problem ----
    def set_problem(self, M, cams6, pts3, feats, cam_idx, pt_idx, cam_id=None, calib: Calib | None = None):
        cams6 = as_f64(cams6, (-1, 6)); pts3 = as_f64(pts3, (-1, 3)); feats = as_f64(feats, (-1, M))
        cam_idx = as_i32(cam_idx); pt_idx = as_i32(pt_idx)
        cam_id = None if cam_id is None else as_i32(cam_id)
        calib = calib or default_calib(self.lib)
        self._check(self.lib.uba_set_problem(self._h, M, cams6.shape[0], pts3.shape[0], feats.shape[0], dptr(cams6), dptr(pts3),
                                             dptr(feats), i32ptr(cam_idx), i32ptr(pt_idx), i32ptr(cam_id), C.byref(calib)))
        self.M = M
        self.n_windows, self.n_cams, self.n_pts, self.n_obs = 1, cams6.shape[0], pts3.shape[0], feats.shape[0]
        self.win_cam_off = np.array([0, self.n_cams], np.int32)

    def set_batch(self, M, win_cam_off, win_pt_off, win_obs_off, cams6, pts3, feats, cam_idx, pt_idx, cam_id=None,
                  calib: Calib | None = None):
        wc = as_i32(win_cam_off); wp = as_i32(win_pt_off); wo = np.ascontiguousarray(win_obs_off, dtype=np.int64)
        cams6 = as_f64(cams6, (-1, 6)); pts3 = as_f64(pts3, (-1, 3)); feats = as_f64(feats, (-1, M))
        cam_idx = as_i32(cam_idx); pt_idx = as_i32(pt_idx)
        cam_id = None if cam_id is None else as_i32(cam_id)
        calib = calib or default_calib(self.lib)
        self._check(self.lib.uba_set_batch(self._h, M, len(wc) - 1, i32ptr(wc), i32ptr(wp), i64ptr(wo), dptr(cams6), dptr(pts3),
                                           dptr(feats), i32ptr(cam_idx), i32ptr(pt_idx), i32ptr(cam_id), C.byref(calib)))
        self.M = M
        self.n_windows, self.n_cams, self.n_pts, self.n_obs = len(wc) - 1, cams6.shape[0], pts3.shape[0], feats.shape[0]
        self.win_cam_off = wc

    def window_advance(self, n_drop, new_cams6, new_pts3, feats, cam_idx, pt_idx, new_pt_cam_id=None, cams6_all=None, pts3_all=None):
        """uba_window_advance; returns pt_id_map (new id of every old point, -1 for erased tracks)."""
        M = self.M
        new_cams6 = as_f64(new_cams6, (-1, 6)); new_pts3 = as_f64(new_pts3, (-1, 3)); feats = as_f64(feats, (-1, M))
        cam_idx = as_i32(cam_idx); pt_idx = as_i32(pt_idx)
        cid = None if new_pt_cam_id is None else as_i32(new_pt_cam_id)
        ca = None if cams6_all is None else as_f64(cams6_all, (-1, 6)); pa = None if pts3_all is None else as_f64(pts3_all, (-1, 3))
        id_map = np.zeros(max(self.n_pts, 1), np.int32)
        self._check(self.lib.uba_window_advance(self._h, n_drop, new_cams6.shape[0], dptr(new_cams6), new_pts3.shape[0], dptr(new_pts3),
                                                i32ptr(cid), feats.shape[0], dptr(feats), i32ptr(cam_idx), i32ptr(pt_idx), dptr(ca), dptr(pa),
                                                i32ptr(id_map)))
        id_map = id_map[:self.n_pts]
        nc = C.c_int(0); npt = C.c_int(0); no = C.c_int64(0); nw = C.c_int(0)
        self._check(self.lib.uba_get_sizes(self._h, C.byref(nw), C.byref(nc), C.byref(npt), C.byref(no)))
        self.n_cams, self.n_pts, self.n_obs = nc.value, npt.value, int(no.value)
        self.win_cam_off = np.array([0, self.n_cams], np.int32)
        return id_map

    # ---- hot path ----
    def optimise(self, fixed_frames: int, check: bool = True):
        sums = (Summary * self.n_windows)()
        rc = self.lib.uba_optimise(self._h, fixed_frames, sums)
        if check:
            self._check(rc, ok=(UBA_OK, UBA_ERR_NUMERICAL, UBA_ERR_INFEASIBLE))
        return rc, list(sums)

    def n_free(self, fixed_frames: int) -> np.ndarray:
        fc = self.tables(fixed_frames)["free_cam"]
        return np.array([(fc[self.win_cam_off[w]:self.win_cam_off[w + 1]] >= 0).sum() for w in range(self.n_windows)])

    def linearize(self, fixed_frames: int, radius: float, want=("residuals", "weights", "cost", "grad_cams", "grad_pts", "B", "C",
                                                                 "W", "S", "rhs", "lm_diag_cams", "lm_diag_pts")):
        M = self.M
        nf = self.n_free(fixed_frames)
        n = 6 * nf
        shapes = {"residuals": (self.n_obs, M), "weights": (self.n_obs,), "cost": (self.n_windows,), "grad_cams": (self.n_cams, 6),
                  "grad_pts": (self.n_pts, 3), "B": (self.n_cams, 6, 6), "C": (self.n_pts, 3, 3), "W": (self.n_obs, 6, 3),
                  "S": (int((n * n).sum()),), "rhs": (int(n.sum()),), "lm_diag_cams": (self.n_cams, 6), "lm_diag_pts": (self.n_pts, 3)}
        out = LinearizationOut()
        arrays = {}
        for k in want:
            arrays[k] = np.zeros(shapes[k], np.float64)
            setattr(out, k, dptr(arrays[k]))
        self._check(self.lib.uba_linearize(self._h, fixed_frames, radius, C.byref(out)))
        if "S" in arrays and self.n_windows == 1:
            arrays["S"] = arrays["S"].reshape(int(n[0]), int(n[0]))
        return arrays

    # ---- results ----
    def cameras(self, out: np.ndarray | None = None) -> np.ndarray:
        """Poses [n_cams, 6]; `out` (C-contiguous float64 of that shape) is filled in place when given."""
        a = np.zeros((self.n_cams, 6)) if out is None else out
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.n_cams, 6)
        self._check(self.lib.uba_get_cameras(self._h, dptr(a)))
        return a

    def points(self, out: np.ndarray | None = None) -> np.ndarray:
        """Points [n_pts, 3] in the caller's order; `out` is filled in place when given."""
        a = np.zeros((self.n_pts, 3)) if out is None else out
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.n_pts, 3)
        self._check(self.lib.uba_get_points(self._h, dptr(a)))
        return a

    def pose_covariances(self) -> np.ndarray:
        a = np.zeros((self.n_cams, 6, 6))
        self._check(self.lib.uba_get_pose_covariances(self._h, dptr(a)))
        return a

    def iterations(self, window: int = 0):
        n = C.c_int(0)
        self._check(self.lib.uba_get_iterations(self._h, window, None, 0, C.byref(n)))
        recs = (Iteration * max(n.value, 1))()
        self._check(self.lib.uba_get_iterations(self._h, window, recs, n.value, C.byref(n)))
        return [{f: getattr(recs[i], f) for f, _ in Iteration._fields_ if f != "pad_"} for i in range(n.value)]

    def tables(self, fixed_frames: int):
        obs_order = np.zeros(self.n_obs, np.int32); pt_off = np.zeros(self.n_pts + 1, np.int64)
        pt_order = np.zeros(self.n_pts, np.int32); free_cam = np.zeros(self.n_cams, np.int32)
        self._check(self.lib.uba_get_tables(self._h, fixed_frames, i32ptr(obs_order), i64ptr(pt_off), i32ptr(pt_order), i32ptr(free_cam)))
        return {"obs_order": obs_order, "pt_obs_off": pt_off, "pt_order": pt_order, "free_cam": free_cam}

    # ---- timing ----
    def set_profiling(self, on: bool):
        self._check(self.lib.uba_set_profiling(self._h, int(on)))

    def timing(self, reset=False) -> dict:
        t = Timing()
        self._check(self.lib.uba_get_timing(self._h, C.byref(t), int(reset)))
        return {n: getattr(t, n) for n, _ in Timing._fields_}

    def time_linearize(self, fixed_frames, radius=1e4, repeats=10, flush_l2=True) -> float:
        ms = C.c_double(0)
        self._check(self.lib.uba_time_linearize(self._h, fixed_frames, radius, repeats, int(flush_l2), C.byref(ms)))
        return ms.value

    def time_iteration(self, fixed_frames, iterations=10, flush_l2=True) -> float:
        ms = C.c_double(0)
        self._check(self.lib.uba_time_iteration(self._h, fixed_frames, iterations, int(flush_l2), C.byref(ms)))
        return ms.value

    def probe_fp64_tflops(self) -> float:
        v = C.c_double(0)
        self._check(self.lib.uba_probe_fp64_tflops(self._h, C.byref(v)))
        return v.value

    # ---- pose-only mode (stereo visual odometry) ----
    def vo_set_matches(self, params: "VoParams", quads) -> np.ndarray:
        q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 8)
        self._check(self.lib.uba_vo_set_matches(self._h, C.byref(params), q.shape[0], q.ctypes.data_as(C.POINTER(C.c_float))))
        self.vo_n = q.shape[0]
        return q

    def vo_points(self) -> np.ndarray:
        a = np.zeros((self.vo_n, 4))
        self._check(self.lib.uba_vo_get_points(self._h, dptr(a)))
        return a

    def vo_linearize(self, state, selection):
        state = as_f64(state); sel = as_i32(selection); n = len(sel)
        A = np.zeros((6, 6)); B = np.zeros(6); res = np.zeros((n, 4)); J = np.zeros((6, 4 * n))
        self._check(self.lib.uba_vo_linearize(self._h, dptr(state), n, i32ptr(sel), dptr(A), dptr(B), dptr(res), dptr(J)))
        return dict(A=A, B=B, res=res, J=J)

    def vo_ransac(self, init, triples):
        init = as_f64(init); tr = as_i32(triples).reshape(-1, 3); nh = tr.shape[0]
        best = C.c_int32(-1); cnt = np.zeros(nh, np.int32); ok = np.zeros(nh, np.int32); st = np.zeros((nh, 6))
        self._check(self.lib.uba_vo_ransac(self._h, dptr(init), nh, i32ptr(tr), C.byref(best), i32ptr(cnt), i32ptr(ok), dptr(st)))
        return dict(best=best.value, counts=cnt, ok=ok, states=st)

    def vo_inliers(self) -> np.ndarray:
        n = C.c_int32(0); idx = np.zeros(max(self.vo_n, 1), np.int32)
        self._check(self.lib.uba_vo_get_inliers(self._h, i32ptr(idx), C.byref(n)))
        return idx[:n.value].copy()

    def vo_refine(self, init, selection=None):
        init = as_f64(init); out = np.zeros(6); conv = C.c_int32(0); it = C.c_int32(0)
        sel = None if selection is None else as_i32(selection)
        self._check(self.lib.uba_vo_refine(self._h, dptr(init), 0 if sel is None else len(sel), i32ptr(sel), dptr(out), C.byref(conv), C.byref(it)))
        return bool(conv.value), out, it.value

    # ---- multi-GPU ----
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
        self._check(self.lib.uba_comm_unique_id(self._h, buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = C.create_string_buffer(bytes(unique_id), NCCL_UNIQUE_ID_BYTES)
        self._check(self.lib.uba_comm_init(self._h, buf, rank, n_ranks))
