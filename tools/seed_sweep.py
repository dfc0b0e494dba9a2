"""Dev tool: step time of the headline configuration for different batch seeds (the multi-GPU bench gives
every rank its own seed and reports the slowest rank)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
dev = torch.device("cuda", 0)
B = 100000
ws = cb.Workspace(B, dev)
for seed in range(42, 50):
    d = synth.make_batch(B, 8, 4, noise=1.0, seed=seed)
    K = torch.from_numpy(d["K"]).to(dev)
    args = dict(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev),
                line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    out = None
    for _ in range(2):
        out = cb.solve_batched(K, **args, workspace=ws, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        out = cb.solve_batched(K, **args, workspace=ws, out=out, timing=True)
    e.record()
    torch.cuda.synchronize()
    t = cb.last_kernel_times()
    it = out.iters.cpu().numpy()
    print(f"seed {seed}: {s.elapsed_time(e) / 5:.2f} ms  K1 {t['solve_fused_kernel']:.2f} K2 {t['straggler_kernel']:.2f}  "
          f"iters mean {it.mean():.1f} max {it.max()} top5 {np.sort(it)[-5:]}", flush=True)
