"""Drop-in scalar API: the reference's pnp / pnl / pnpl signatures
(cvxpnpl.py:523-530, 555-562, 586-595) and its K-first plugin classes
(benchmarks/toolkit/methods/pnp.py:85-93, pnl.py:37-48, pnpl.py:49-58).

Each call runs the CUDA path on a batch of one and converts the per-problem
status back into the reference's behaviour: a list of (R, t) numpy pairs of
length 1, 2 or 4; NaN pose on solver NaN (cvxpnpl.py:493-498);
numpy.linalg.LinAlgError / NotImplementedError where the reference raises them
(cvxpnpl.py:212, 341); a warning when the solution is not certifiably optimal
(cvxpnpl.py:516-519).
"""
import warnings
from typing import List, Tuple

import numpy as np
import torch

from . import batched as _b

Pose = Tuple[np.ndarray, np.ndarray]


def _unpack(res: _b.BatchedPoses, verbose: bool) -> List[Pose]:
    status = int(res.status[0].item())
    code = status & _b.ST_CODE_MASK
    if code == _b.ST_NAN:
        if verbose:
            warnings.warn("The SDP solver did not return a valid solution. Increasing max_iters might solve the issue.")
        return [(np.full((3, 3), np.nan), np.full(3, np.nan))]
    if code == _b.ST_SINGULAR:
        raise np.linalg.LinAlgError("Singular matrix")
    if code == _b.ST_RANK0:
        raise NotImplementedError
    n = int(res.n_poses[0].item())
    R = res.R[0, :n].cpu().numpy()
    t = res.t[0, :n].cpu().numpy()
    if status & _b.FLAG_NOT_CERTIFIED:
        warnings.warn("The solution is not certifiably optimal.")
    if verbose:
        print(f"[cvxpnpl_b200] iters={int(res.iters[0])} status={code} n_poses={n} "
              f"obj={res.obj[0].tolist() if res.obj is not None else None}")
    return [(R[i].copy(), t[i].copy()) for i in range(n)]


def _one(x, tail):
    x = torch.as_tensor(np.asarray(x, dtype=np.float64) if not isinstance(x, torch.Tensor) else x)
    return x.reshape((1, -1) + tail)


def pnp(pts_2d, pts_3d, K, eps: float = 1e-9, max_iters: int = 2500, verbose: bool = False) -> List[Pose]:
    """Compute object poses from point 2D-3D correspondences (cvxpnpl.py:523-552).

    pts_2d -- n x 2 pixels; pts_3d -- n x 3 points; K -- 3 x 3 intrinsics."""
    res = _b.solve_batched(np.asarray(K, dtype=np.float64), pts_2d=_one(pts_2d, (2,)), pts_3d=_one(pts_3d, (3,)),
                           eps=eps, max_iters=max_iters)
    return _unpack(res, verbose)


def pnl(line_2d, line_3d, K, eps: float = 1e-9, max_iters: int = 2500, verbose: bool = False) -> List[Pose]:
    """Compute object poses from line 2D-3D correspondences (cvxpnpl.py:555-583).

    line_2d -- n x 2 x 2 (line, endpoint, xy); line_3d -- n x 2 x 3."""
    res = _b.solve_batched(np.asarray(K, dtype=np.float64), line_2d=_one(line_2d, (2, 2)),
                           line_3d=_one(line_3d, (2, 3)), eps=eps, max_iters=max_iters)
    return _unpack(res, verbose)


def pnpl(pts_2d, line_2d, pts_3d, line_3d, K, eps: float = 1e-9, max_iters: int = 2500,
         verbose: bool = False) -> List[Pose]:
    """Compute object poses from point and line correspondences (cvxpnpl.py:586-627)."""
    kw = {}
    if np.size(pts_2d):
        kw.update(pts_2d=_one(pts_2d, (2,)), pts_3d=_one(pts_3d, (3,)))
    if np.size(line_2d):
        kw.update(line_2d=_one(line_2d, (2, 2)), line_3d=_one(line_3d, (2, 3)))
    res = _b.solve_batched(np.asarray(K, dtype=np.float64), eps=eps, max_iters=max_iters, **kw)
    return _unpack(res, verbose)


def _nan_pose() -> List[Pose]:
    return [(np.full((3, 3), np.nan), np.full(3, np.nan))]


class CvxPnPL:
    """Plugin class with the reference's K-first `estimate_pose` shape so it can be
    handed to the reference's `Suite(methods=[...])`.  The three reference classes
    (methods/pnp.py:85-93, pnl.py:37-48, pnpl.py:49-58) differ only in which
    keyword arguments they take; one class accepts all of them."""

    name = "CvxPnPL"
    loaded = True

    @staticmethod
    def estimate_pose(K, pts_2d=None, line_2d=None, pts_3d=None, line_3d=None):
        n_p = 0 if pts_2d is None else len(pts_2d)
        n_l = 0 if line_2d is None else len(line_2d)
        # requires a minimum of 3 elements (methods/pnpl.py:55-57)
        if n_p + n_l < 3:
            return _nan_pose()
        if n_l == 0:
            return pnp(pts_2d, pts_3d, K)
        if n_p == 0:
            return pnl(line_2d, line_3d, K)
        return pnpl(pts_2d, line_2d, pts_3d, line_3d, K)


def _rc_unpack(res, verbose):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # _solve_relaxation_rc (rc.py:68-122) has no optimality warning
        return _unpack(res, verbose)


def rc(pts_2d, pts_3d, K, eps: float = 1e-9, max_iters: int = 2500, verbose: bool = False) -> List[Pose]:
    """The "rc" ablation for points (benchmarks/toolkit/methods/pnp.py:58-82 -> rc.py:68-122): pnp with
    the redundant row-orthonormality equalities removed from the SDP (16 instead of 22 equalities)."""
    res = _b.solve_batched(np.asarray(K, dtype=np.float64), pts_2d=_one(pts_2d, (2,)), pts_3d=_one(pts_3d, (3,)),
                           eps=eps, max_iters=max_iters, variant="rc")
    return _rc_unpack(res, verbose)


rc_pnp = rc


def rc_pnl(line_2d, line_3d, K, eps: float = 1e-9, max_iters: int = 2500, verbose: bool = False) -> List[Pose]:
    """The "rc" ablation for lines (benchmarks/toolkit/methods/pnl.py:11-34 -> rc.py:68-122)."""
    res = _b.solve_batched(np.asarray(K, dtype=np.float64), line_2d=_one(line_2d, (2, 2)),
                           line_3d=_one(line_3d, (2, 3)), eps=eps, max_iters=max_iters, variant="rc")
    return _rc_unpack(res, verbose)


def rc_pnpl(pts_2d, line_2d, pts_3d, line_3d, K, eps: float = 1e-9, max_iters: int = 2500,
            verbose: bool = False) -> List[Pose]:
    """The "rc" ablation for points and lines (benchmarks/toolkit/methods/pnpl.py:12-46 -> rc.py:68-122)."""
    kw = {}
    if np.size(pts_2d):
        kw.update(pts_2d=_one(pts_2d, (2,)), pts_3d=_one(pts_3d, (3,)))
    if np.size(line_2d):
        kw.update(line_2d=_one(line_2d, (2, 2)), line_3d=_one(line_3d, (2, 3)))
    res = _b.solve_batched(np.asarray(K, dtype=np.float64), eps=eps, max_iters=max_iters, variant="rc", **kw)
    return _rc_unpack(res, verbose)


def null(pts_2d, pts_3d, K) -> List[Pose]:
    """The "null" ablation (benchmarks/toolkit/methods/pnp.py:24-55): no SDP, the smallest right singular
    vector of A projected onto SO(3)."""
    res = _b.null_batched(_one(pts_2d, (2,)), _one(pts_3d, (3,)), np.asarray(K, dtype=np.float64))
    return [(res.R[0, 0].cpu().numpy().copy(), res.t[0, 0].cpu().numpy().copy())]
