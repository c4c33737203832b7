"""ctypes loader for the C-ABI device backend (include/spn_b200.h).

There is no CPU fallback: if the sm_100a library has not been built, importing the
compute entry points fails loudly instead of routing anywhere else.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
# SPN_B200_LIB points tuning runs at an alternative build of the same library
KERNEL_LIB = os.environ.get("SPN_B200_LIB") or os.path.join(LIB_DIR, "libspn_b200.so")
COMM_LIB = os.path.join(LIB_DIR, "libspn_b200_comm.so")
HOST_LIB = os.path.join(LIB_DIR, "libspn_host.so")


class SpnError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        msg = error_string(status)
        super().__init__(f"{where}: {msg} (status {status})" if where else f"{msg} (status {status})")


_P = c_void_p
_IP = c_void_p  # int* on the device or host, passed as an address

# name -> argtypes, for every `int spn_*` entry point of include/spn_b200.h that lives
# in libspn_b200.so (the comm section lives in libspn_b200_comm.so, see COMM_SIGNATURES)
KERNEL_SIGNATURES = {
    "spn_version": [],
    "spn_device_count": [POINTER(c_int)],
    "spn_set_device": [c_int],
    "spn_get_device": [POINTER(c_int)],
    "spn_device_sm_count": [POINTER(c_int)],
    "spn_device_sync": [],
    "spn_stream_create": [POINTER(c_void_p)],
    "spn_stream_create_priority": [POINTER(c_void_p), c_int],
    "spn_stream_destroy": [_P],
    "spn_stream_sync": [_P],
    "spn_malloc": [POINTER(c_void_p), c_size_t],
    "spn_free": [_P],
    "spn_host_malloc": [POINTER(c_void_p), c_size_t],
    "spn_host_free": [_P],
    "spn_memset": [_P, c_int, c_size_t, _P],
    "spn_memcpy_h2d": [_P, _P, c_size_t, _P, c_int],
    "spn_memcpy_d2h": [_P, _P, c_size_t, _P, c_int],
    "spn_memcpy_d2d": [_P, _P, c_size_t, _P, c_int],
    "spn_event_create": [POINTER(c_void_p)],
    "spn_event_destroy": [_P],
    "spn_event_record": [_P, _P],
    "spn_event_sync": [_P],
    "spn_stream_wait_event": [_P, _P],
    "spn_event_elapsed_ms": [_P, _P, POINTER(c_float)],
    "spn_graph_begin_capture": [_P],
    "spn_graph_end_capture": [_P, POINTER(c_void_p)],
    "spn_graph_launch": [_P, _P],
    "spn_graph_destroy": [_P],
    "spn_bs0": [c_int, _P, _P, _P],
    "spn_copy_f64": [c_int64, _P, _P, _P],
    "spn_axpby_f64": [c_int64, c_double, _P, c_double, _P, _P],
    "spn_norm2_f64": [c_int64, _P, _P, _P, _P],
    "spn_dot_f64": [c_int64, _P, _P, _P, _P, _P],
    "spn_cg_update_f64": [c_int64, _P, c_double, _P, _P, _P, _P],
    "spn_bs3_1": [c_int, c_int, _P, _P, _P],
    "spn_bs3_2": [c_int, _P, _P, _P],
    "spn_bs4_1": [c_int, c_int, _P, _P, _P, _P],
    "spn_bs4_2": [c_int, _P, _P, _P],
    "spn_bs5_1": [c_int, c_int, _P, c_double, _P, _P, _P],
    "spn_bs5_2": [c_int, _P, _P, _P],
    "spn_ogs_nodes_per_block": [],
    "spn_ogs_gather": [c_int, c_int, _IP, _IP, _IP, _P, _P, c_int, c_int, _P],
    "spn_ogs_scatter": [c_int, c_int, _IP, _IP, _IP, _P, _P, c_int, _P],
    "spn_ogs_gather_scatter": [c_int, c_int, _IP, _IP, _IP, _IP, _IP, _P, c_int, c_int, _P],
    "spn_ogs_extract": [c_int, c_int, _IP, _P, _P, c_int, _P],
    "spn_ogs_plan_create": [POINTER(c_void_p), c_int, c_int, _IP, _IP, _IP, _IP],
    "spn_ogs_plan_destroy": [_P],
    "spn_ogs_plan_info": [_P, POINTER(c_int), POINTER(c_int), POINTER(c_size_t)],
    "spn_ogs_plan_gather": [_P, c_int, _P, _P, c_int, c_int, _P],
    "spn_ogs_plan_scatter": [_P, c_int, _P, _P, c_int, _P],
    "spn_ogs_plan_gather_scatter": [_P, c_int, _P, c_int, c_int, _P],
    "spn_ogs_plan_gather_scatter_to": [_P, c_int, _P, _P, c_int, c_int, _P],
    "spn_debug_set": [c_int, c_int],
    "spn_ogs_halo_gather_pack": [c_int, c_int, c_int, _IP, _IP, _IP, _P, _P, _P, c_int, c_int, _P],
    "spn_ogs_halo_combine_scatter": [c_int, c_int, c_int, _IP, _IP, _P, _P, _IP, _IP, _P, c_int, c_int, _P],
    # peer-memory halo exchange (spn_p2p_plan_t / spn_p2p_state_t live in device memory)
    "spn_ipc_get_handle": [_P, _P],
    "spn_ipc_open_handle": [_P, POINTER(c_void_p)],
    "spn_ipc_close_handle": [_P],
    "spn_ogs_halo_gather_pack_p2p": [c_int, c_int, _IP, _IP, _IP, _IP, _P, _P, _P, _P, c_int, c_int, _P],
    "spn_ogs_extract_p2p": [c_int, c_int, _IP, _IP, _P, _P, _P, c_int, _P],
    "spn_ogs_halo_combine_scatter_p2p": [c_int, c_int, c_int, _IP, _IP, _P, _P, _P, _IP, _IP, _P, c_int, c_int, _P],
}

COMM_SIGNATURES = {
    "spn_comm_unique_id": [_P],
    "spn_comm_init": [POINTER(c_void_p), _P, c_int, c_int],
    "spn_comm_destroy": [_P],
    "spn_comm_rank": [_P, POINTER(c_int)],
    "spn_comm_size": [_P, POINTER(c_int)],
    "spn_comm_exchange": [_P, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P],
    "spn_comm_allreduce_sum_f64": [_P, _P, c_int, _P],
    "spn_comm_barrier": [_P, _P],
}

_kernel_lib = None
_comm_lib = None


def _bind(lib, signatures):
    for name, argtypes in signatures.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export the ABI
        fn.argtypes = argtypes
        fn.restype = c_int
    return lib


def kernel_lib():
    """libspn_b200.so; raises if it has not been built (no fallback)."""
    global _kernel_lib
    if _kernel_lib is None:
        if not os.path.exists(KERNEL_LIB):
            raise ImportError(
                f"{KERNEL_LIB} is missing: build it with `make` (or __graft_entry__.build()). "
                "streamparanumal_b200 has no CPU fallback.")
        lib = ctypes.CDLL(KERNEL_LIB, mode=ctypes.RTLD_GLOBAL)
        _bind(lib, KERNEL_SIGNATURES)
        lib.spn_error_string.argtypes = [c_int]
        lib.spn_error_string.restype = c_char_p
        lib.spn_reduce_workspace_bytes.argtypes = []
        lib.spn_reduce_workspace_bytes.restype = c_size_t
        _kernel_lib = lib
    return _kernel_lib


def comm_lib():
    """libspn_b200_comm.so (NCCL); load torch first so its bundled libnccl.so.2 is the one bound."""
    global _comm_lib
    if _comm_lib is None:
        if not os.path.exists(COMM_LIB):
            raise ImportError(f"{COMM_LIB} is missing: build it with `make`.")
        kernel_lib()
        try:
            import torch  # noqa: F401  (loads libnccl.so.2 by soname before our DT_NEEDED lookup)
            import torch.distributed  # noqa: F401
        except Exception:
            pass
        _comm_lib = _bind(ctypes.CDLL(COMM_LIB, mode=ctypes.RTLD_GLOBAL), COMM_SIGNATURES)
    return _comm_lib


def error_string(status):
    try:
        return kernel_lib().spn_error_string(int(status)).decode()
    except Exception:
        return f"spn status {status}"


def check(status, where=""):
    if status != 0:
        raise SpnError(status, where)
