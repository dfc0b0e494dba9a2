"""Dev tool (torchrun, N ranks): per-step times of the sharded (strong-scaling) solve of one 1e5 batch."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
from cvxpnpl_b200.distributed import RecordGatherer, shard_bounds, solve_sharded
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 100000
d0 = synth.make_batch(B, 8, 4, noise=1.0, seed=42)
full = {k: torch.from_numpy(d0[k]).to(dev) for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")}
K = torch.from_numpy(d0["K"]).to(dev)
lo, hi = shard_bounds(B, rank, world)
gat = RecordGatherer(B, dev); ws = cb.Workspace(hi - lo, dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for rep in range(6):
    for _ in range(3):
        solve_sharded(K, **full, gatherer=gat, workspace=ws); flush.zero_()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for i in range(10):          # like bench.py: no host synchronisation between the steps
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        solve_sharded(K, **full, gatherer=gat, workspace=ws)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ts = [s.elapsed_time(e) for s, e in evs]
    print(f"rank {rank} rep {rep}: mean {np.mean(ts):.2f} max {max(ts):.2f} ms  {[round(t, 1) for t in ts]}", flush=True)
dist.destroy_process_group()
