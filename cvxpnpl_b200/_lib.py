"""ctypes binding of the C-ABI library (include/cvxpnpl_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) as
cvxpnpl_b200/libcvxpnpl_b200.so.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present every entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CVXPNPL_B200_LIB: developer override (A/B builds of the same C ABI)
LIB_PATH = os.environ.get("CVXPNPL_B200_LIB") or os.path.join(_HERE, "libcvxpnpl_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class Desc(ctypes.Structure):
    """Mirror of `cvxpnpl_b200_desc` (include/cvxpnpl_b200.h)."""

    _fields_ = [
        ("batch", ctypes.c_int64),
        ("n_pts", ctypes.c_int32),
        ("n_lines", ctypes.c_int32),
        ("k_batched", ctypes.c_int32),
        ("anderson", ctypes.c_int32),
        ("K", ctypes.c_void_p),
        ("pts_2d", ctypes.c_void_p),
        ("pts_3d", ctypes.c_void_p),
        ("line_2d", ctypes.c_void_p),
        ("line_3d", ctypes.c_void_p),
        ("eps", ctypes.c_double),
        ("max_iters", ctypes.c_int32),
        ("sweeps", ctypes.c_int32),
        ("variant", ctypes.c_int32),
        ("handoff", ctypes.c_int32),
        ("rho_rel", ctypes.c_double),
        ("alpha", ctypes.c_double),
        ("sigma", ctypes.c_double),
        ("R", ctypes.c_void_p),
        ("t", ctypes.c_void_p),
        ("n_poses", ctypes.c_void_p),
        ("status", ctypes.c_void_p),
        ("iters", ctypes.c_void_p),
        ("obj", ctypes.c_void_p),
        ("Z", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("fp32_iters", ctypes.c_int32),
        ("timing", ctypes.c_int32),
        ("skip_prepass", ctypes.c_int32),
        ("psd_mode", ctypes.c_int32),
        ("record", ctypes.c_void_p),
    ]


EXPORTS = (
    "cvxpnpl_b200_version",
    "cvxpnpl_b200_last_error",
    "cvxpnpl_b200_workspace_bytes",
    "cvxpnpl_b200_solve",
    "cvxpnpl_b200_assemble",
    "cvxpnpl_b200_solve_sdp",
    "cvxpnpl_b200_extract",
    "cvxpnpl_b200_last_launch_count",
    "cvxpnpl_b200_fp64_probe",
    "cvxpnpl_b200_null",
    "cvxpnpl_b200_kernel_times",
    "cvxpnpl_b200_prepass",
)

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises LibraryMissing if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  cvxpnpl_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.cvxpnpl_b200_version.restype = ctypes.c_char_p
    lib.cvxpnpl_b200_last_error.restype = ctypes.c_char_p
    lib.cvxpnpl_b200_workspace_bytes.restype = ctypes.c_size_t
    lib.cvxpnpl_b200_workspace_bytes.argtypes = [ctypes.c_int64]
    lib.cvxpnpl_b200_solve.restype = ctypes.c_int
    lib.cvxpnpl_b200_solve.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p]
    lib.cvxpnpl_b200_assemble.restype = ctypes.c_int
    lib.cvxpnpl_b200_assemble.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.cvxpnpl_b200_solve_sdp.restype = ctypes.c_int
    lib.cvxpnpl_b200_solve_sdp.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_void_p]
    lib.cvxpnpl_b200_extract.restype = ctypes.c_int
    lib.cvxpnpl_b200_extract.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p]
    lib.cvxpnpl_b200_null.restype = ctypes.c_int
    lib.cvxpnpl_b200_null.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p]
    lib.cvxpnpl_b200_prepass.restype = ctypes.c_int
    lib.cvxpnpl_b200_prepass.argtypes = [ctypes.POINTER(Desc), ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
    lib.cvxpnpl_b200_kernel_times.restype = ctypes.c_int
    lib.cvxpnpl_b200_kernel_times.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.cvxpnpl_b200_last_launch_count.restype = ctypes.c_int
    lib.cvxpnpl_b200_fp64_probe.restype = ctypes.c_int
    lib.cvxpnpl_b200_fp64_probe.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().cvxpnpl_b200_last_error().decode()
        raise RuntimeError(f"cvxpnpl_b200 call failed (code {rc}): {msg}")
