"""The C-ABI library loads on a CPU-only box and exports every symbol include/uba.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

from uasl_motion_estimation_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "uba.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uba_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert sorted(capi.EXPORTED_SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol(gpu_lib):
    for name in declared_symbols():
        assert hasattr(gpu_lib, name), name
    assert gpu_lib.uba_version() == 100


def test_config_defaults_are_the_reference_values(gpu_lib):
    cfg = capi.default_config(gpu_lib)
    assert cfg.loss_kind == capi.LOSS_HUBER and cfg.loss_scale == 1.0      # BundleAdjuster.h:397,:447
    assert cfg.function_tolerance == 1e-3 and cfg.max_solver_time_s == 1.0  # :417-419
    assert cfg.max_iterations == 50 and cfg.initial_radius == 1e4 and cfg.min_relative_decrease == 1e-3
    assert cfg.jacobi_scaling == 1 and cfg.use_bounds == 1 and cfg.compute_covariance == 0


def test_no_cpu_fallback_create_fails_without_a_gpu(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    cfg = capi.default_config(gpu_lib)
    rc = gpu_lib.uba_create(C.byref(cfg), C.byref(h))
    assert rc == capi.UBA_ERR_CUDA and not h.value
    assert b"no CPU path" in gpu_lib.uba_last_error(None) or b"CUDA" in gpu_lib.uba_last_error(None)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        capi.load(tmp_path / "libuba.so")


def test_product_sources_never_touch_the_oracle():
    for p in list((ROOT / "uasl_motion_estimation_b200").rglob("*")):
        if p.suffix in {".py", ".cu", ".cpp", ".h"}:
            t = p.read_text()
            assert "uba_oracle" not in t and "oracle_binding" not in t and "uba_ref_" not in t, p
