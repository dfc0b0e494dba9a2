"""Batched entry points: B independent problems held as torch CUDA tensors.

These are thin wrappers that allocate outputs, fill the C descriptor
(include/cvxpnpl_b200.h) with raw device pointers and launch on torch's current
CUDA stream.  torch is plumbing (device memory + streams) only.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

ST_OK, ST_MAX_ITERS, ST_NAN, ST_SINGULAR, ST_RANK0 = 0, 1, 2, 3, 4
#: correspondences per problem from which the streaming (bandwidth-bound) assembly is used
LARGE_N = 256
ST_CODE_MASK = 0xFF
FLAG_NOT_CERTIFIED = 0x100


@dataclass
class BatchedPoses:
    """Result of a batched solve.  Candidate poses are NaN padded to 4."""

    R: torch.Tensor        # [B, 4, 3, 3] world->camera
    t: torch.Tensor        # [B, 4, 3]
    n_poses: torch.Tensor  # [B] int32
    status: torch.Tensor   # [B] int32 (code | flags)
    iters: torch.Tensor    # [B] int32
    obj: Optional[torch.Tensor] = None  # [B, 2] (r'Qr of candidate 0, dual objective)
    Z: Optional[torch.Tensor] = None    # [B, 10, 10]
    launches: int = 0
    record: Optional[torch.Tensor] = None   # [B, 15] packed (R0 | t0 | n_poses | status | iters), see distributed.RECORD


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("cvxpnpl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _dev_f64(x, device, shape_tail, name):
    x = torch.as_tensor(x)
    if x.device.type != "cuda":
        x = x.to(device)
    x = x.to(torch.float64).contiguous()
    if tuple(x.shape[-len(shape_tail):]) != tuple(shape_tail) if shape_tail else False:
        raise ValueError(f"{name}: expected trailing shape {shape_tail}, got {tuple(x.shape)}")
    return x


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


class Workspace:
    """Caller-owned scratch (device) reused across calls of the same batch size."""

    def __init__(self, batch, device):
        lib = _lib.load()
        device = torch.device(device)
        with torch.cuda.device(device):   # the size depends on the SM count of the device it is for
            self.nbytes = int(lib.cvxpnpl_b200_workspace_bytes(batch))
        self.buf = torch.empty(max(self.nbytes // 8, 1), dtype=torch.float64, device=device)


def solve_batched(K, pts_2d=None, pts_3d=None, line_2d=None, line_3d=None, eps=1e-9, max_iters=2500,
                  sweeps=0, rho_rel=0.0, alpha=0.0, sigma=0.0, anderson=True, variant="full", return_Z=False, return_obj=True, workspace=None,
                  out: Optional[BatchedPoses] = None, device=None, handoff=0, admm_dtype="f64",
                  fp32_iters=400, timing=False, _prepass_hook=None, record=None, psd="track") -> BatchedPoses:
    """Solve B problems.  pts_2d [B,n,2], pts_3d [B,n,3], line_2d [B,m,2,2],
    line_3d [B,m,2,3], K [3,3] or [B,3,3]; any of the point / line pairs may be
    omitted (PnP / PnL / PnPL: cvxpnpl.py:523-627).  variant="rc" solves the ablation
    of benchmarks/toolkit/methods/rc.py (row-orthonormality equalities removed).
    handoff: passes after which a problem still iterating once the work queue is empty
    moves to the warp-per-problem straggler kernel (0 = default, < 0 = never).
    admm_dtype="f32": "fp32 ADMM + fp64 extraction" (BASELINE.json configs[3]) -- the
    iterations that bring a problem into the linear tail run in FP32 (at most
    fp32_iters), the FP64 solver finishes to `eps`.
    psd: PSD projection of the ADMM iteration -- "track" (default): two tracked eigenpairs refined once per
    iteration with a certificate, problems that fail it are finished with the full decomposition; "full": a full
    10x10 eigen-decomposition every iteration; "track1": the tracked projection with one thread per problem instead
    of two (kept for A/B measurements).
    record: optional [B,15] float64 CUDA tensor (may be a slice of an all-gather buffer) that the finish
    kernel fills with the packed row (R0 | t0 | n_poses | status | iters) of every problem."""
    _require_cuda()
    lib = _lib.load()
    if device is None:
        for x in (pts_2d, line_2d, K):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                device = x.device
                break
        else:
            device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    def _has(x):
        if x is None:
            return False
        x = torch.as_tensor(x)
        return x.dim() >= 2 and x.shape[1] > 0

    have_p, have_l = _has(pts_2d), _has(line_2d)
    if not (have_p or have_l):
        raise ValueError("need at least one point or line correspondence")
    B = None
    if have_p:
        pts_2d = _dev_f64(pts_2d, device, (2,), "pts_2d")
        pts_3d = _dev_f64(pts_3d, device, (3,), "pts_3d")
        if pts_2d.dim() != 3 or pts_3d.dim() != 3 or pts_2d.shape[:2] != pts_3d.shape[:2]:
            raise ValueError("pts_2d must be [B,n,2] and pts_3d [B,n,3]")
        B = pts_2d.shape[0]
    if have_l:
        line_2d = _dev_f64(line_2d, device, (2, 2), "line_2d")
        line_3d = _dev_f64(line_3d, device, (2, 3), "line_3d")
        if line_2d.dim() != 4 or line_3d.dim() != 4 or line_2d.shape[:2] != line_3d.shape[:2]:
            raise ValueError("line_2d must be [B,m,2,2] and line_3d [B,m,2,3]")
        if B is not None and line_2d.shape[0] != B:
            raise ValueError("points and lines disagree on the batch size")
        B = line_2d.shape[0]
    K = _dev_f64(K, device, (3, 3), "K")
    k_batched = K.dim() == 3
    if k_batched and K.shape[0] != B:
        raise ValueError("batched K must be [B,3,3]")

    if variant not in ("full", "rc"):
        raise ValueError("variant must be 'full' or 'rc'")
    if admm_dtype not in ("f64", "f32"):
        raise ValueError("admm_dtype must be 'f64' or 'f32'")
    if psd not in ("track", "full", "track1"):
        raise ValueError("psd must be 'track', 'full' or 'track1'")
    n_corr = (pts_2d.shape[1] if have_p else 0) + (line_2d.shape[1] if have_l else 0)
    if n_corr >= LARGE_N and B > 0:
        # many correspondences per problem (benchmarks/scalability/pnp.py:37-40): the
        # assembly is a bandwidth-bound streaming reduction with its own kernels; the
        # SDP and the extraction then run as stages on the [B,9,9] / [B,10,10] matrices.
        # Options that only exist on the fused path are refused, never silently dropped.
        unsupported = [name for name, given in (("admm_dtype='f32'", admm_dtype != "f64"), ("out", out is not None),
                                                ("handoff", handoff != 0), ("timing", bool(timing)),
                                                ("record", record is not None),
                                                ("_prepass_hook", _prepass_hook is not None)) if given]
        if unsupported:
            raise NotImplementedError(f"{', '.join(unsupported)}: not available with >= {LARGE_N} correspondences per "
                                      "problem (streaming assembly + stage kernels)")
        with torch.cuda.device(device):
            Q, Bm = assemble_batched(K, pts_2d if have_p else None, pts_3d if have_p else None,
                                     line_2d if have_l else None, line_3d if have_l else None)
            launches = int(lib.cvxpnpl_b200_last_launch_count())
            Z, dobj, iters, status = solve_sdp_batched(Q, eps=eps, max_iters=max_iters, sweeps=sweeps, rho_rel=rho_rel,
                                                       alpha=alpha, sigma=sigma, anderson=anderson, variant=variant,
                                                       n_pts=pts_2d.shape[1] if have_p else 0)
            launches += int(lib.cvxpnpl_b200_last_launch_count())
            res = extract_batched(Z, Q, Bm, dobj, eps=eps)
            launches += int(lib.cvxpnpl_b200_last_launch_count())
            res.iters = iters
            # keep the solver's status (MAX_ITERS / NaN) where extraction itself succeeded
            res.status = torch.where((res.status & ST_CODE_MASK) == ST_OK, res.status | status, res.status)
            res.Z = Z if return_Z else None
            if not return_obj:
                res.obj = None
            res.launches = launches
        return res
    with torch.cuda.device(device):
        if out is None:
            out = BatchedPoses(
                R=torch.empty((B, 4, 3, 3), dtype=torch.float64, device=device),
                t=torch.empty((B, 4, 3), dtype=torch.float64, device=device),
                n_poses=torch.empty(B, dtype=torch.int32, device=device),
                status=torch.empty(B, dtype=torch.int32, device=device),
                iters=torch.empty(B, dtype=torch.int32, device=device),
                obj=torch.empty((B, 2), dtype=torch.float64, device=device) if return_obj else None,
                Z=torch.empty((B, 10, 10), dtype=torch.float64, device=device) if return_Z else None,
            )
        if B == 0:
            return out
        if workspace is None:
            workspace = Workspace(B, device)
        d = _lib.Desc()
        d.batch = B
        d.n_pts = pts_2d.shape[1] if have_p else 0
        d.n_lines = line_2d.shape[1] if have_l else 0
        d.k_batched = int(k_batched)
        d.K = _ptr(K)
        d.pts_2d, d.pts_3d = _ptr(pts_2d if have_p else None), _ptr(pts_3d if have_p else None)
        d.line_2d, d.line_3d = _ptr(line_2d if have_l else None), _ptr(line_3d if have_l else None)
        d.eps, d.max_iters, d.sweeps, d.rho_rel, d.alpha = float(eps), int(max_iters), int(sweeps), float(rho_rel), float(alpha)
        d.sigma = float(sigma)
        d.anderson = 0 if anderson else -1
        d.variant = {"full": 0, "rc": 1}[variant]
        d.handoff = int(handoff)
        d.fp32_iters = int(fp32_iters) if admm_dtype == "f32" else 0
        d.timing = int(bool(timing))
        d.psd_mode = {"track": 0, "full": 1, "track1": 2}[psd]
        d.R, d.t, d.n_poses, d.status, d.iters = _ptr(out.R), _ptr(out.t), _ptr(out.n_poses), _ptr(out.status), _ptr(out.iters)
        d.obj, d.Z = _ptr(out.obj), _ptr(out.Z)
        if record is not None:
            if (not record.is_cuda or record.dtype != torch.float64 or tuple(record.shape) != (B, 15)
                    or not record.is_contiguous()):
                raise ValueError("record must be a contiguous [B,15] float64 CUDA tensor")
            out.record = record
        d.record = _ptr(record)
        d.workspace, d.workspace_bytes = _ptr(workspace.buf), workspace.nbytes
        stream = torch.cuda.current_stream(device).cuda_stream
        extra = 0
        if _prepass_hook is not None:
            # the caller streams the inputs in and runs the pre-pass chunk by chunk (pipeline.solve_from_host)
            extra = _prepass_hook(lib, d, stream)
            d.skip_prepass = 1
        _lib.check(lib.cvxpnpl_b200_solve(ctypes.byref(d), ctypes.c_void_p(stream)))
        out.launches = int(lib.cvxpnpl_b200_last_launch_count()) + extra
    return out


def pnp_batched(pts_2d, pts_3d, K, **kw) -> BatchedPoses:
    """Batched cvxpnpl.pnp (cvxpnpl.py:523-552), same argument order."""
    return solve_batched(K, pts_2d=pts_2d, pts_3d=pts_3d, **kw)


def pnl_batched(line_2d, line_3d, K, **kw) -> BatchedPoses:
    """Batched cvxpnpl.pnl (cvxpnpl.py:555-583), same argument order."""
    return solve_batched(K, line_2d=line_2d, line_3d=line_3d, **kw)


def pnpl_batched(pts_2d, line_2d, pts_3d, line_3d, K, **kw) -> BatchedPoses:
    """Batched cvxpnpl.pnpl (cvxpnpl.py:586-627), same argument order."""
    return solve_batched(K, pts_2d=pts_2d, pts_3d=pts_3d, line_2d=line_2d, line_3d=line_3d, **kw)


# ---- stage entry points (parity tests of the individual reference functions) ----


def assemble_batched(K, pts_2d=None, pts_3d=None, line_2d=None, line_3d=None, staging="tma"):
    """-> Q [B,9,9] (= A'A, cvxpnpl.py:475) and Bmat [B,3,9] (cvxpnpl.py:623).  With >= 256 correspondences per problem
    the assembly is a streaming reduction; staging="tma" (default) stages the point slabs through shared memory with
    bulk-asynchronous copies, "loads" keeps the plain-load kernel (A/B measurements)."""
    _require_cuda()
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device())
    have_p = pts_2d is not None and torch.as_tensor(pts_2d).numel() > 0
    have_l = line_2d is not None and torch.as_tensor(line_2d).numel() > 0
    d = _lib.Desc()
    if have_p:
        pts_2d, pts_3d = _dev_f64(pts_2d, device, (2,), "pts_2d"), _dev_f64(pts_3d, device, (3,), "pts_3d")
        B, d.n_pts = pts_2d.shape[0], pts_2d.shape[1]
    if have_l:
        line_2d, line_3d = _dev_f64(line_2d, device, (2, 2), "line_2d"), _dev_f64(line_3d, device, (2, 3), "line_3d")
        B, d.n_lines = line_2d.shape[0], line_2d.shape[1]
    K = _dev_f64(K, device, (3, 3), "K")
    d.batch, d.k_batched, d.K = B, int(K.dim() == 3), _ptr(K)
    d.pts_2d, d.pts_3d = _ptr(pts_2d if have_p else None), _ptr(pts_3d if have_p else None)
    d.line_2d, d.line_3d = _ptr(line_2d if have_l else None), _ptr(line_3d if have_l else None)
    if staging not in ("tma", "loads"):
        raise ValueError("staging must be 'tma' or 'loads'")
    d.psd_mode = 0 if staging == "tma" else 1
    Q = torch.empty((B, 9, 9), dtype=torch.float64, device=device)
    Bm = torch.empty((B, 3, 9), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    _lib.check(lib.cvxpnpl_b200_assemble(ctypes.byref(d), _ptr(Q), _ptr(Bm), ctypes.c_void_p(stream)))
    return Q, Bm


def solve_sdp_batched(Q, eps=1e-9, max_iters=2500, sweeps=0, rho_rel=0.0, alpha=0.0, sigma=0.0, anderson=True,
                      n_pts=8, variant="full"):
    """Q [B,9,9] -> (Z [B,10,10], dobj [B], iters [B], status [B]); the scs.solve
    call of cvxpnpl.py:478-492 (variant="rc": of benchmarks/toolkit/methods/rc.py:90-96).
    n_pts (points behind each Q) only selects the default rho / sigma where they are left 0."""
    _require_cuda()
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device())
    Q = _dev_f64(Q, device, (9, 9), "Q")
    B = Q.shape[0]
    Z = torch.empty((B, 10, 10), dtype=torch.float64, device=device)
    obj = torch.empty((B, 2), dtype=torch.float64, device=device)
    iters = torch.empty(B, dtype=torch.int32, device=device)
    status = torch.empty(B, dtype=torch.int32, device=device)
    ws = Workspace(B, device)
    d = _lib.Desc()
    d.batch = B
    d.n_pts = int(n_pts)
    d.eps, d.max_iters, d.sweeps, d.rho_rel, d.alpha = float(eps), int(max_iters), int(sweeps), float(rho_rel), float(alpha)
    d.sigma = float(sigma)
    d.anderson = 0 if anderson else -1
    d.variant = {"full": 0, "rc": 1}[variant]
    d.Z, d.obj, d.iters, d.status = _ptr(Z), _ptr(obj), _ptr(iters), _ptr(status)
    d.workspace, d.workspace_bytes = _ptr(ws.buf), ws.nbytes
    stream = torch.cuda.current_stream(device).cuda_stream
    _lib.check(lib.cvxpnpl_b200_solve_sdp(ctypes.byref(d), _ptr(Q), ctypes.c_void_p(stream)))
    return Z, obj[:, 1].contiguous(), iters, status


def extract_batched(Z, Q, Bmat, dobj=None, eps=1e-9) -> BatchedPoses:
    """Z [B,10,10], Q [B,9,9], Bmat [B,3,9] -> poses; cvxpnpl.py:493-520."""
    _require_cuda()
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device())
    Z, Q, Bmat = _dev_f64(Z, device, (10, 10), "Z"), _dev_f64(Q, device, (9, 9), "Q"), _dev_f64(Bmat, device, (3, 9), "Bmat")
    B = Z.shape[0]
    out = BatchedPoses(
        R=torch.empty((B, 4, 3, 3), dtype=torch.float64, device=device),
        t=torch.empty((B, 4, 3), dtype=torch.float64, device=device),
        n_poses=torch.empty(B, dtype=torch.int32, device=device),
        status=torch.empty(B, dtype=torch.int32, device=device),
        iters=torch.zeros(B, dtype=torch.int32, device=device),
        obj=torch.empty((B, 2), dtype=torch.float64, device=device),
    )
    d = _lib.Desc()
    d.batch, d.eps = B, float(eps)
    d.R, d.t, d.n_poses, d.status, d.obj = _ptr(out.R), _ptr(out.t), _ptr(out.n_poses), _ptr(out.status), _ptr(out.obj)
    if dobj is not None:
        dobj = _dev_f64(dobj, device, (), "dobj")
    stream = torch.cuda.current_stream(device).cuda_stream
    _lib.check(lib.cvxpnpl_b200_extract(ctypes.byref(d), _ptr(Z), _ptr(Q), _ptr(Bmat), _ptr(dobj), ctypes.c_void_p(stream)))
    return out


def null_batched(pts_2d, pts_3d, K) -> BatchedPoses:
    """Batched "null" baseline (benchmarks/toolkit/methods/pnp.py:24-55): no SDP, the
    smallest right singular vector of A projected onto SO(3)."""
    _require_cuda()
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device())
    pts_2d, pts_3d = _dev_f64(pts_2d, device, (2,), "pts_2d"), _dev_f64(pts_3d, device, (3,), "pts_3d")
    K = _dev_f64(K, device, (3, 3), "K")
    B = pts_2d.shape[0]
    out = BatchedPoses(
        R=torch.empty((B, 4, 3, 3), dtype=torch.float64, device=device),
        t=torch.empty((B, 4, 3), dtype=torch.float64, device=device),
        n_poses=torch.empty(B, dtype=torch.int32, device=device),
        status=torch.empty(B, dtype=torch.int32, device=device),
        iters=torch.zeros(B, dtype=torch.int32, device=device),
    )
    d = _lib.Desc()
    d.batch, d.n_pts, d.n_lines, d.k_batched = B, pts_2d.shape[1], 0, int(K.dim() == 3)
    d.K, d.pts_2d, d.pts_3d = _ptr(K), _ptr(pts_2d), _ptr(pts_3d)
    d.R, d.t, d.n_poses, d.status, d.iters = _ptr(out.R), _ptr(out.t), _ptr(out.n_poses), _ptr(out.status), _ptr(out.iters)
    stream = torch.cuda.current_stream(device).cuda_stream
    _lib.check(lib.cvxpnpl_b200_null(ctypes.byref(d), ctypes.c_void_p(stream)))
    out.launches = int(lib.cvxpnpl_b200_last_launch_count())
    return out


def measure_fp64_peak(device=None, iters=4000, repeats=5):
    """Measured FP64 FMA throughput (TFLOP/s) of the device: best of `repeats`
    launches of the library's probe kernel, timed with CUDA events."""
    _require_cuda()
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count
        out = torch.empty(n_sm * 8 * 256, dtype=torch.float64, device=device)
        flops = ctypes.c_int64(0)
        stream = torch.cuda.current_stream(device).cuda_stream
        best = float("inf")
        for _ in range(repeats + 1):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.check(lib.cvxpnpl_b200_fp64_probe(_ptr(out), out.numel(), int(iters), ctypes.byref(flops),
                                                   ctypes.c_void_p(stream)))
            e.record()
            torch.cuda.synchronize(device)
            best = min(best, s.elapsed_time(e))
    return flops.value / (best * 1e-3) / 1e12


KERNEL_NAMES = ("pre_kernel", "admm32_kernel", "ortho_kernel", "solve_fused_kernel", "straggler_kernel",
                "solve_fused_kernel<resume>", "finish_kernel", "solve_track_kernel", "redecomp_kernel")


def last_kernel_times():
    """Device time (ms) of each kernel of the last `solve_batched(..., timing=True)` on this
    host thread (CUDA events on the launching stream, recorded by the library)."""
    lib = _lib.load()
    ms = (ctypes.c_float * len(KERNEL_NAMES))()
    _lib.check(lib.cvxpnpl_b200_kernel_times(ms, len(KERNEL_NAMES)))
    return dict(zip(KERNEL_NAMES, [float(x) for x in ms]))
