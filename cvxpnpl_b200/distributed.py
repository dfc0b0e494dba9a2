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
    (one NCCL call) does the job.  (Convenience form that copies; RecordGatherer is the
    copy-free one the bench and solve_sharded use.)"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    g = RecordGatherer(total, local.device, group=group)
    g.slot[: local.shape[0]].copy_(local)
    return g.gather()


class RecordGatherer:
    """The path's one collective without staging copies: a persistent [world * width, RECORD] buffer
    whose slice `slot` ([width, RECORD], this rank's place) is handed to the solver as its packed-record
    output (`solve_batched(record=gatherer.local)`), so the finish kernel writes the rows where the
    all-gather reads them and `all_gather_into_tensor` runs IN PLACE (input = the rank's slice of the
    output: no pad, no concatenate).  width = the largest shard; with a batch that does not divide evenly
    the trailing row of the smaller shards is padding and `gather()` drops it."""

    def __init__(self, total: int, device, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total = total
        self.bounds = [shard_bounds(total, r, self.world) for r in range(self.world)]
        self.width = max(hi - lo for lo, hi in self.bounds) if total else 0
        self.buf = torch.zeros((self.world * self.width, RECORD), dtype=torch.float64, device=device)
        self.slot = self.buf[self.rank * self.width: (self.rank + 1) * self.width]
        lo, hi = self.bounds[self.rank]
        self.local = self.slot[: hi - lo]      # [own shard, RECORD]: pass this as `record=`
        self.even = all(hi - lo == self.width for lo, hi in self.bounds)

    def gather(self) -> torch.Tensor:
        """-> [total, RECORD] on every rank (a view of the buffer when the shards are equal)."""
        if self.world > 1 and self.width:
            dist.all_gather_into_tensor(self.buf, self.slot, group=self.group)
        if self.even:
            return self.buf
        return torch.cat([self.buf[r * self.width: r * self.width + (hi - lo)] for r, (lo, hi) in enumerate(self.bounds)],
                         dim=0)


def solve_sharded(K, pts_2d=None, pts_3d=None, line_2d=None, line_3d=None, group=None, solver=None, gatherer=None,
                  **kw):
    """Every rank passes the FULL batch (host or device tensors); each solves its own
    contiguous shard on its GPU and the poses are all-gathered (strong scaling: SURVEY 8e,
    "rank g gets problems [g B/G, (g+1) B/G)").  Returns (R [B,3,3], t [B,3], n_poses, status,
    iters) of candidate 0 on every rank.  `solver` defaults to cvxpnpl_b200.solve_batched
    (injectable for CPU tests); `gatherer` (a RecordGatherer for this batch size) can be kept
    across calls so nothing is allocated per step."""
    ref = pts_2d if pts_2d is not None else line_2d
    total = ref.shape[0]
    real = solver is None
    if real:
        from .batched import solve_batched as solver
    if gatherer is None:
        dev = ref.device if (ref.is_cuda or not real) else torch.device("cuda", torch.cuda.current_device())
        gatherer = RecordGatherer(total, dev, group=group)
    lo, hi = gatherer.bounds[gatherer.rank]
    sl = lambda x: None if x is None else x[lo:hi]  # noqa: E731
    Kl = K[lo:hi] if (hasattr(K, "dim") and K.dim() == 3) else K
    if real:
        # the finish kernel writes the packed rows straight into this rank's slice of the gather buffer
        solver(Kl, pts_2d=sl(pts_2d), pts_3d=sl(pts_3d), line_2d=sl(line_2d), line_3d=sl(line_3d),
               record=gatherer.local, **kw)
    else:
        res = solver(Kl, pts_2d=sl(pts_2d), pts_3d=sl(pts_3d), line_2d=sl(line_2d), line_3d=sl(line_3d), **kw)
        gatherer.local.copy_(pack_record(res.R[:, 0], res.t[:, 0], res.n_poses, res.status, res.iters))
    return unpack_record(gatherer.gather())
