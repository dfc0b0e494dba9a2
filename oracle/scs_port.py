"""oracle/scs_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/scs_port.c, the CPU restatement of the conic solve the
reference hands to third-party SCS (cvxpnpl.py:485-489).  See the header of
scs_port.c for what is restated and for the "parity unpinned" note.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libscs_port.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "scs_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-o", _SO, src, "-lm"]
        )
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.scs_port_setup.argtypes = [dp, dp, ctypes.c_int]
        lib.scs_port_setup.restype = ctypes.c_int
        lib.scs_port_solve.argtypes = [dp, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_double, dp, dp, dp, dp]
        lib.scs_port_solve.restype = ctypes.c_int
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


_setup_key = None


def solve(A, b, c, eps_abs=1e-9, eps_rel=0.0, max_iters=2500, alpha=1.5, cscale=10.0):
    """min c'x s.t. Ax + s = b, s in {0}^nz x S_+^10 (A dense (nz+55) x 55; nz = 22, or 16 for
    the "rc" ablation of benchmarks/toolkit/methods/rc.py)."""
    global _setup_key
    lib = _load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    m = A.shape[0]
    assert A.shape == (m, 55) and b.shape == (m,) and c.shape == (55,) and m in (77, 71)
    key = (A.tobytes(), b.tobytes())
    if _setup_key != key:
        assert lib.scs_port_setup(_p(A), _p(b), m - 55) == 0
        _setup_key = key
    x = np.empty(55)
    y = np.empty(77)
    s = np.empty(77)
    info = np.empty(8)
    lib.scs_port_solve(_p(c), eps_abs, eps_rel, int(max_iters), alpha, cscale, _p(x), _p(y), _p(s), _p(info))
    status = {1: "solved", 2: "solved_inaccurate", -1: "infeasible_or_unbounded"}[int(info[6])]
    y, s = y[:m], s[:m]
    return {
        "x": x, "y": y, "s": s,
        "info": {"pobj": info[0], "dobj": info[1], "res_pri": info[2], "res_dual": info[3],
                 "gap": info[4], "iter": int(info[5]), "status": status},
    }
