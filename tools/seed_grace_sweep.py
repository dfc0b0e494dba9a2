"""Dev tool: step time of the headline configuration over batch seeds x hand-over grace periods (the
multi-GPU bench gives every rank its own seed and reports the slowest rank, so the worst seed matters)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
dev = torch.device("cuda", 0)
B = 100000
graces = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "10,20,40").split(",")]
seeds = range(42, 42 + (int(sys.argv[2]) if len(sys.argv) > 2 else 12))
ws = cb.Workspace(B, dev)
tab = {g: [] for g in graces}
for seed in seeds:
    d = synth.make_batch(B, 8, 4, noise=1.0, seed=seed)
    K = torch.from_numpy(d["K"]).to(dev)
    args = dict(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev),
                line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    row = []
    for g in graces:
        out = None
        for _ in range(2):
            out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=g)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(4):
            out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=g)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 4
        tab[g].append(ms)
        row.append(f"g{g} {ms:.2f}")
    it = out.iters.cpu().numpy()
    print(f"seed {seed}: {'  '.join(row)}  iters max {it.max()} top3 {np.sort(it)[-3:]}", flush=True)
for g in graces:
    a = np.array(tab[g])
    print(f"grace {g}: mean {a.mean():.2f} max {a.max():.2f} ms")
