"""Builds and binds tests/host/host_harness.cpp (test-only host compilation of the
device routines; see the header of that file)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
_SO = os.path.join(_HERE, "_build", "libhost_harness.so")
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lib = None


def _p(a):
    return a.ctypes.data_as(_dp)


def load():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "host_harness.cpp")] + [
            os.path.join(_ROOT, "cvxpnpl_b200", "csrc", f) for f in ("pnpl_core.cuh", "pnpl_dr.inl", "pnpl_extract.cuh", "pnpl_solve.cuh", "pnpl_track.cuh", "pnpl_track2.cuh")]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _SO, srcs[0]])
        _lib = ctypes.CDLL(_SO)
    return _lib


def load_variant(tag, defines):
    """A second build of the harness with extra -D flags (algorithm variants compared against the default build);
    returns a context manager that routes the functions of this module to it."""
    import contextlib
    so = os.path.join(_HERE, "_build", f"libhost_{tag}.so")
    srcs = [os.path.join(_HERE, "host_harness.cpp")] + [
        os.path.join(_ROOT, "cvxpnpl_b200", "csrc", f) for f in ("pnpl_core.cuh", "pnpl_dr.inl", "pnpl_extract.cuh", "pnpl_solve.cuh", "pnpl_track.cuh", "pnpl_track2.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread"] + [f"-D{x}" for x in defines] + ["-o", so, srcs[0]])
    variant = ctypes.CDLL(so)

    @contextlib.contextmanager
    def use():
        global _lib
        load()
        keep, _lib_new = _lib, variant
        _set(_lib_new)
        try:
            yield
        finally:
            _set(keep)
    return use


def _set(lib):
    global _lib
    _lib = lib


def solve(d, eps=1e-9, max_iters=2500, sweeps=0, rho_rel=0.0, alpha=0.0, sigma=0.0, anderson=1, variant=0, fp32_iters=0):
    lib = load()
    B, n_pts, n_lines = d["pts_2d"].shape[0], d["pts_2d"].shape[1], d["line_2d"].shape[1]
    c = {k: np.ascontiguousarray(d[k], dtype=np.float64) for k in ("K", "pts_2d", "pts_3d", "line_2d", "line_3d")}
    R, t = np.empty((B, 4, 3, 3)), np.empty((B, 4, 3))
    n, st, it = (np.empty(B, np.int32) for _ in range(3))
    obj, Z = np.empty((B, 2)), np.empty((B, 10, 10))
    lib.host_solve(ctypes.c_int64(B), n_pts, n_lines, _p(c["K"]), int(c["K"].ndim == 3), _p(c["pts_2d"]),
                   _p(c["pts_3d"]), _p(c["line_2d"]), _p(c["line_3d"]), ctypes.c_double(eps), max_iters, sweeps,
                   ctypes.c_double(rho_rel), ctypes.c_double(alpha), ctypes.c_double(sigma), int(anderson), int(variant), int(fp32_iters), _p(R), _p(t), n.ctypes.data_as(_ip),
                   st.ctypes.data_as(_ip), it.ctypes.data_as(_ip), _p(obj), _p(Z))
    return dict(R=R, t=t, n_poses=n, status=st, iters=it, obj=obj, Z=Z)


def solve_track(d, eps=1e-9, max_iters=2500, rho_rel=0.0, alpha=0.0, sigma=0.0, anderson=1, variant=0, two=False):
    """The tracked solver (pnpl_track.cuh) chained like the CUDA kernels chain it; r["fallbacks"][b] = 1 where
    the certificate failed and the full-decomposition path finished the problem.  two=True: the role-split form
    (pnpl_track2.cuh), two host threads per problem with a barrier where the kernel has its named barriers."""
    lib = load()
    B, n_pts, n_lines = d["pts_2d"].shape[0], d["pts_2d"].shape[1], d["line_2d"].shape[1]
    c = {k: np.ascontiguousarray(d[k], dtype=np.float64) for k in ("K", "pts_2d", "pts_3d", "line_2d", "line_3d")}
    R, t = np.empty((B, 4, 3, 3)), np.empty((B, 4, 3))
    n, st, it, fb = (np.empty(B, np.int32) for _ in range(4))
    obj, Z = np.empty((B, 2)), np.empty((B, 10, 10))
    (lib.host_solve_track2 if two else lib.host_solve_track)(ctypes.c_int64(B), n_pts, n_lines, _p(c["K"]), int(c["K"].ndim == 3), _p(c["pts_2d"]),
                         _p(c["pts_3d"]), _p(c["line_2d"]), _p(c["line_3d"]), ctypes.c_double(eps), max_iters,
                         ctypes.c_double(rho_rel), ctypes.c_double(alpha), ctypes.c_double(sigma), int(anderson),
                         int(variant), _p(R), _p(t), n.ctypes.data_as(_ip), st.ctypes.data_as(_ip),
                         it.ctypes.data_as(_ip), _p(obj), _p(Z), fb.ctypes.data_as(_ip))
    return dict(R=R, t=t, n_poses=n, status=st, iters=it, obj=obj, Z=Z, fallbacks=fb)


def extract(Z, Q, Bm):
    lib = load()
    Z, Q, Bm = (np.ascontiguousarray(x, dtype=np.float64) for x in (Z, Q, Bm))
    R, t, st = np.empty((4, 3, 3)), np.empty((4, 3)), np.zeros(1, np.int32)
    n = lib.host_extract(_p(Z), _p(Q), _p(Bm), _p(R), _p(t), st.ctypes.data_as(_ip))
    return n, R, t, int(st[0])


def quartic(c):
    lib = load()
    c = np.ascontiguousarray(c, dtype=np.float64)
    x = np.empty(4)
    n = lib.host_quartic(_p(c), _p(x))
    return x[:n]


def plateau_update(plat, res2, res2_prev):
    """plateau_update of pnpl_solve.cuh on one (state, residual) step: returns (new state, jump length)."""
    lib = load()
    st = ctypes.c_int32(plat)
    tau = lib.host_plateau_update(ctypes.byref(st), ctypes.c_double(res2), ctypes.c_double(res2_prev))
    return st.value, tau
