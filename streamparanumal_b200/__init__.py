"""streamparanumal_b200 — B200-native (sm_100a) device backend for streamParanumal's
hot path (BS0–BS8), behind the C ABI in include/spn_b200.h.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (libspn_b200.so, libspn_b200_comm.so)
  host/      C++ host layer mirroring the reference's platform/ogs/mesh API + BS executables
  kernels.py kernel-level Python surface (ctypes over the C ABI; torch only as memory handle)
"""
from . import _lib  # noqa: F401
from ._lib import SpnError  # noqa: F401

__all__ = ["SpnError", "kernels"]


def __getattr__(name):
    if name == "kernels":
        import importlib
        return importlib.import_module(".kernels", __name__)
    raise AttributeError(name)
