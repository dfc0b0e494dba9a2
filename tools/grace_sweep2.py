"""Dev tool: hand-over grace sweep for the tracked solver (headline batch and PnP-8)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
dev = torch.device("cuda", 0)
B = 100000
for n_pts, n_lines in ((8, 4), (8, 0), (0, 6)):
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=42)
    K = torch.from_numpy(d["K"]).to(dev)
    args = {}
    if n_pts:
        args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
    if n_lines:
        args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    ws = cb.Workspace(B, dev)
    for grace in (24, 32, 40, 48, 56, 64, 80, 100):
        if grace < 0 and n_lines == 6:
            continue
        out = None
        for _ in range(2):
            out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=grace)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=grace, timing=True)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = {k.replace("_kernel", ""): round(v, 3) for k, v in cb.last_kernel_times().items() if v > 0}
        print(f"{n_pts}+{n_lines} grace {grace:4d}: {np.median(ts):.3f} ms {t} handed_back {int(ws.buf[:16].view(torch.int64)[6].item())}", flush=True)
