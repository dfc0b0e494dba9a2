"""Multi-GPU plumbing: one process per GPU (torchrun), batches of independent
problems sharded by contiguous ranges, ONE all-gather of the output poses.

The SDP solves are independent, so there is no data-path
collective inside the solve; the only exchange is collecting
[R (9) | t (3) | n_poses | status | iters] per problem onto every rank.
Works with the NCCL backend (GPU tensors) and with gloo (CPU tensors; used by
the CPU tests of this module's logic).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

#: columns of the gathered record
RECORD = 15  # 9 (R of candidate 0, row-major) + 3 (t) + n_poses + status + iters


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of the batch owned by `rank`: rank g gets problems
    [g*B/G, (g+1)*B/G) with the remainder spread over the first ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_record(R0: torch.Tensor, t0: torch.Tensor, n_poses: torch.Tensor, status: torch.Tensor,
                iters: torch.Tensor) -> torch.Tensor:
    """[b, 15] float64 record of the first candidate pose and the per-problem flags."""
    b = R0.shape[0]
    return torch.cat([R0.reshape(b, 9), t0.reshape(b, 3), n_poses.to(torch.float64)[:, None],
                      status.to(torch.float64)[:, None], iters.to(torch.float64)[:, None]], dim=1).contiguous()


def unpack_record(rec: torch.Tensor):
    R = rec[:, :9].reshape(-1, 3, 3)
    t = rec[:, 9:12]
    return R, t, rec[:, 12].to(torch.int32), rec[:, 13].to(torch.int32), rec[:, 14].to(torch.int32)


def gather_records(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather ragged shards (sizes given by shard_bounds) into [total, RECORD].
    Shards are padded to the largest shard so a single all_gather_into_tensor
    (one NCCL call) does the job."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def solve_sharded(K, pts_2d=None, pts_3d=None, line_2d=None, line_3d=None, group=None, solver=None, **kw):
    """Every rank passes the FULL batch (host or device tensors); each solves its own
    contiguous shard on its GPU and the poses are all-gathered.  Returns
    (R [B,3,3], t [B,3], n_poses, status, iters) of candidate 0 on every rank.
    `solver` defaults to cvxpnpl_b200.solve_batched (injectable for CPU tests)."""
    if solver is None:
        from .batched import solve_batched as solver
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    ref = pts_2d if pts_2d is not None else line_2d
    total = ref.shape[0]
    lo, hi = shard_bounds(total, rank, world)
    sl = lambda x: None if x is None else x[lo:hi]  # noqa: E731
    Kl = K[lo:hi] if (hasattr(K, "dim") and K.dim() == 3) else K
    res = solver(Kl, pts_2d=sl(pts_2d), pts_3d=sl(pts_3d), line_2d=sl(line_2d), line_3d=sl(line_3d), **kw)
    rec = pack_record(res.R[:, 0], res.t[:, 0], res.n_poses, res.status, res.iters)
    return unpack_record(gather_records(rec, total, group))
