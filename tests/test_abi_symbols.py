"""The C-ABI libraries load without a GPU and export every entry point that
include/spn_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    names = declared("spn_b200.h")
    for must in ("spn_bs0", "spn_copy_f64", "spn_axpby_f64", "spn_norm2_f64", "spn_dot_f64",
                 "spn_cg_update_f64", "spn_ogs_gather", "spn_ogs_scatter", "spn_ogs_gather_scatter",
                 "spn_ogs_extract", "spn_comm_exchange", "spn_comm_allreduce_sum_f64"):
        assert must in names


def test_every_declared_symbol_is_exported():
    from streamparanumal_b200 import _lib
    assert os.path.exists(_lib.KERNEL_LIB), "run `make` (or __graft_entry__.build()) first"
    kernel = ctypes.CDLL(_lib.KERNEL_LIB)
    comm_names = [n for n in declared("spn_b200.h") if n.startswith("spn_comm_")]
    for name in declared("spn_b200.h"):
        if name in comm_names:
            continue
        assert hasattr(kernel, name), f"{name} declared in spn_b200.h but not exported by libspn_b200.so"
    # the python binding table covers the same set
    bound = set(_lib.KERNEL_SIGNATURES) | {"spn_error_string", "spn_reduce_workspace_bytes"} | set(_lib.COMM_SIGNATURES)
    assert set(declared("spn_b200.h")) == bound


def test_comm_library_exports():
    from streamparanumal_b200 import _lib
    assert os.path.exists(_lib.COMM_LIB)
    try:
        import torch  # noqa: F401  binds torch's libnccl.so.2 first, as at run time
    except Exception:
        pass
    comm = ctypes.CDLL(_lib.COMM_LIB)
    for name in declared("spn_b200.h"):
        if name.startswith("spn_comm_"):
            assert hasattr(comm, name), name


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from streamparanumal_b200 import _lib
    lib = _lib.kernel_lib()
    assert lib.spn_version() >= 100
    # compute entry points must fail loudly, not compute on the host
    rc = lib.spn_copy_f64(4, 8, 8, None)
    assert rc != 0
    assert "fallback" in _lib.error_string(rc) or "CUDA" in _lib.error_string(rc) or rc > 0
    with pytest.raises(TypeError):
        from streamparanumal_b200 import kernels
        a = torch.zeros(4, dtype=torch.float64)
        kernels.copy(4, a, a.clone())
