"""Batched synthetic benchmark suite: the reference's
SynthSuite (benchmarks/toolkit/suites/synth.py:45-346, suite.py:8-110) with whole
cells of the (n_elements x noise) grid drawn and solved as ONE batch on the GPU.

What is restated: random poses (synth.py:13-42), Kinect intrinsics (49-51),
correspondence generators for points / lines / points+lines (276-346; the PnPL
split is drawn per problem like synth.py:323-324 but, as one batch needs one
shape, per CELL here), pinhole projection (suite.py:17-19), the pose error metric
(suite.py:22-33, angle in degrees) and the multi-pose disambiguation with 20
random support points (suite.py:95-108).  torch is used for RNG / elementwise
plumbing on the device; every pose comes from the CUDA solver.
"""
from dataclasses import dataclass
from typing import Dict, Iterable, Tuple

import math
import torch

from . import synth as _synth
from .batched import solve_batched


def random_poses(B, gen, device):
    """synth.py:13-42 on the device: Rodrigues(axis ~ normalised U(-.5,.5)^3, angle ~ U(0,2pi)),
    t = (U(-.5,.5), U(-.5,.5), U(.6, 2.2))."""
    u = lambda *s: torch.rand(*s, generator=gen, device=device, dtype=torch.float64)  # noqa: E731
    axis = u(B, 3) - 0.5
    axis = axis / axis.norm(dim=1, keepdim=True)
    ang = 2 * math.pi * u(B)
    Kx = torch.zeros((B, 3, 3), dtype=torch.float64, device=device)
    Kx[:, 0, 1], Kx[:, 0, 2] = -axis[:, 2], axis[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = axis[:, 2], -axis[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -axis[:, 1], axis[:, 0]
    eye = torch.eye(3, dtype=torch.float64, device=device)
    R = eye + torch.sin(ang)[:, None, None] * Kx + (1 - torch.cos(ang))[:, None, None] * (Kx @ Kx)
    t = torch.cat([u(B, 2) - 0.5, 1.6 * u(B, 1) + 0.6], dim=1)
    return R, t


def project_points(pts, K, R, t):
    """suite.py:17-19, batched: pts [B,n,3] -> pixels [B,n,2]."""
    pc = (pts @ R.transpose(1, 2) + t[:, None, :]) @ K.T
    return pc[..., :2] / pc[..., 2:]


def generate(B, n_pts, n_lines, noise, gen, device, K=None):
    """One batch of problems on the device (synth.py:276-346)."""
    K = torch.as_tensor(_synth.K_KINECT, dtype=torch.float64, device=device) if K is None else K
    R, t = random_poses(B, gen, device)
    P = _synth.LENGTH * (torch.rand((B, n_pts + 2 * n_lines, 3), generator=gen, device=device,
                                    dtype=torch.float64) - 0.5)
    p2 = project_points(P, K, R, t)
    if noise > 0:
        p2 = p2 + noise * torch.randn(p2.shape, generator=gen, device=device, dtype=torch.float64)
    return {"K": K, "R_gt": R, "t_gt": t,
            "pts_2d": p2[:, :n_pts].contiguous(), "pts_3d": P[:, :n_pts].contiguous(),
            "line_2d": p2[:, n_pts:].reshape(B, n_lines, 2, 2).contiguous(),
            "line_3d": P[:, n_pts:].reshape(B, n_lines, 2, 3).contiguous()}


def rotation_angle_deg(Ra, Rb):
    """Geodesic angle in degrees (suite.py:8-14, 29); atan2 form, see synth.rotation_angle."""
    E = Ra.transpose(-1, -2) @ Rb
    sk = torch.stack([E[..., 2, 1] - E[..., 1, 2], E[..., 0, 2] - E[..., 2, 0], E[..., 1, 0] - E[..., 0, 1]], dim=-1)
    return torch.rad2deg(torch.atan2(sk.norm(dim=-1), E.diagonal(dim1=-2, dim2=-1).sum(-1) - 1.0))


def disambiguate(res, batch, gen, n_support=20):
    """suite.py:95-108: among the candidate poses pick the one with the smallest summed
    reprojection error of `n_support` random support points projected with the ground
    truth.  Returns R [B,3,3], t [B,3] (NaN where the solver produced no pose)."""
    B = res.R.shape[0]
    dev = res.R.device
    S = torch.rand((B, n_support, 3), generator=gen, device=dev, dtype=torch.float64) - 0.5
    ref = project_points(S, batch["K"], batch["R_gt"], batch["t_gt"])
    err = torch.full((B, 4), float("inf"), dtype=torch.float64, device=dev)
    for k in range(4):
        pk = project_points(S, batch["K"], res.R[:, k], res.t[:, k])
        e = (pk - ref).norm(dim=-1).sum(dim=-1)
        valid = (k < res.n_poses) & torch.isfinite(e)
        err[:, k] = torch.where(valid, e, err[:, k])
    best = err.argmin(dim=1)
    idx = torch.arange(B, device=dev)
    R, t = res.R[idx, best].clone(), res.t[idx, best].clone()
    none = ~torch.isfinite(err.min(dim=1).values)
    R[none], t[none] = float("nan"), float("nan")
    return R, t


@dataclass
class CellResult:
    ang_median_deg: float
    trans_median: float
    ang_mean_deg: float
    failed: float          # fraction without a finite pose
    iters_median: float
    multi: float           # fraction with more than one candidate


def run_cell(kind, n_elements, noise, runs, seed=42, device=None) -> CellResult:
    """One cell of the reference's grid (synth.py:225-273): `runs` problems of `kind`
    in {"pnp", "pnl", "pnpl"} with n_elements correspondences and pixel noise sigma."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if kind == "pnp":
        n_pts, n_lines = n_elements, 0
    elif kind == "pnl":
        n_pts, n_lines = 0, n_elements
    else:
        n_pts = int(torch.randint(1, n_elements, (1,), generator=gen, device=device))  # synth.py:323
        n_lines = n_elements - n_pts
    batch = generate(runs, n_pts, n_lines, noise, gen, device)
    res = solve_batched(batch["K"], pts_2d=batch["pts_2d"] if n_pts else None, pts_3d=batch["pts_3d"] if n_pts else None,
                        line_2d=batch["line_2d"] if n_lines else None, line_3d=batch["line_3d"] if n_lines else None)
    R, t = disambiguate(res, batch, gen)
    ang = rotation_angle_deg(batch["R_gt"], R)
    tr = (t - batch["t_gt"]).norm(dim=1) / batch["t_gt"].norm(dim=1)
    ok = torch.isfinite(ang) & torch.isfinite(tr)
    return CellResult(float(ang[ok].median()), float(tr[ok].median()), float(ang[ok].mean()),
                      float((~ok).double().mean()), float(res.iters.double().median()),
                      float((res.n_poses > 1).double().mean()))


def run_grid(kind, n_elements: Iterable[int] = (4, 6, 8, 10, 12), noises: Iterable[float] = (0.0, 1.0, 2.0),
             runs=1000, seed=42, device=None) -> Dict[Tuple[int, float], CellResult]:
    """The grid of benchmarks/synth/p*.py:23 (n in {4,6,8,10,12}, sigma in {0,1,2} px,
    1000 runs per cell by default, suites/__init__.py:29)."""
    return {(n, s): run_cell(kind, n, s, runs, seed + 131 * i + j, device)
            for i, n in enumerate(n_elements) for j, s in enumerate(noises)}
