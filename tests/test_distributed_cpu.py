"""CPU test of the N>1 host logic (gloo, world_size 2): sharding by contiguous
ranges and the single all-gather of the pose records.  The CUDA solve is replaced
by a deterministic stand-in solver, so only the plumbing is exercised here."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_shard_bounds_cover_batch():
    from cvxpnpl_b200.distributed import shard_bounds
    for total in (0, 1, 7, 100_000, 100_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_solver(K, pts_2d=None, pts_3d=None, line_2d=None, line_3d=None, **kw):
    from cvxpnpl_b200.batched import BatchedPoses
    b = pts_2d.shape[0]
    key = pts_2d[:, 0, 0]
    R = torch.full((b, 4, 3, 3), float("nan"), dtype=torch.float64)
    t = torch.full((b, 4, 3), float("nan"), dtype=torch.float64)
    R[:, 0] = key[:, None, None] * torch.ones(3, 3, dtype=torch.float64)
    t[:, 0] = key[:, None] + torch.arange(3, dtype=torch.float64)
    return BatchedPoses(R=R, t=t, n_poses=torch.ones(b, dtype=torch.int32),
                        status=(key.to(torch.int32) % 3), iters=key.to(torch.int32) + 100)


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cvxpnpl_b200.distributed import solve_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts_2d = torch.arange(total, dtype=torch.float64)[:, None, None].repeat(1, 4, 2)
    pts_3d = torch.zeros((total, 4, 3), dtype=torch.float64)
    R, t, n, st, it = solve_sharded(torch.eye(3, dtype=torch.float64), pts_2d=pts_2d, pts_3d=pts_3d,
                                    solver=_fake_solver)
    ok = (R.shape == (total, 3, 3)
          and torch.equal(R[:, 0, 0], torch.arange(total, dtype=torch.float64))
          and torch.equal(t[:, 2], torch.arange(total, dtype=torch.float64) + 2)
          and torch.equal(st, (torch.arange(total) % 3).to(torch.int32))
          and torch.equal(it, (torch.arange(total) + 100).to(torch.int32))
          and bool((n == 1).all()))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gloo_world2_gather(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_host_stager_part_bounds():
    """HostStager(parts=...): sub-batch boundaries are contiguous, monotone and cover the batch for counts and fractions,
    ragged and tiny batches included (pure host logic)."""
    from cvxpnpl_b200.pipeline import part_bounds
    for B in (0, 1, 7, 100000, 30001):
        for fr in ([1, 1, 1], [0.3, 0.7], [0.15, 0.45, 0.4], [1.0], [5, 1, 1, 1, 1, 1]):
            b = part_bounds(B, fr)
            assert b[0] == 0 and b[-1] == B and len(b) == len(fr) + 1
            assert all(x <= y for x, y in zip(b[:-1], b[1:]))
    assert part_bounds(100000, [1, 1, 1]) == [0, 33333, 66667, 100000]
    assert part_bounds(100000, [0.3, 0.7]) == [0, 30000, 100000]
