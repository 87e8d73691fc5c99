"""uasl_motion_estimation_b200 — B200-native windowed bundle-adjustment inner loop.

The product is ``lib/libuba.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/uba.h``); this package is the Python-side plumbing: ctypes binding, the
synthetic generator wrapper, point sharding for multi-GPU runs and a mirror of the
reference's ``me::optimisation::BundleAdjuster<M>`` interface.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
