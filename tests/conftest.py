import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding as ob
    ob.lib()
    return ob


@pytest.fixture(scope="session")
def emu_lib():
    """Serial host emulation of libuba's thread-independent kernels (tests/emu): host-logic tests only."""
    from uasl_motion_estimation_b200 import capi
    d = ROOT / "tests" / "emu"
    subprocess.run(["make", "-s", "-C", str(d)], check=True, capture_output=True)
    # the pose-only VO kernels are shared-memory kernels: not part of the emulation build
    return capi.load(d / "libuba_emu.so", allow_missing=tuple(n for n in capi.EXPORTED_SYMBOLS if n.startswith("uba_vo_")))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library.  Fails (does not skip) when it is missing: there is no fallback."""
    from uasl_motion_estimation_b200 import capi
    return capi.default_lib()
