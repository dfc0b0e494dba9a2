"""Dev tool: time the batched solve for the three BASELINE configurations over the
hand-over grace period and the ADMM precision (CUDA events, 5 calls each)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth

dev = torch.device("cuda", 0)
graces = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "-1,20,40,80".split(","))]
for (npt, nl) in ((8, 4), (8, 0), (0, 6)):
    B = 100000
    d = synth.make_batch(B, npt, nl, noise=1.0, seed=42)
    K = torch.from_numpy(d["K"]).to(dev)
    args = {}
    if npt:
        args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
    if nl:
        args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    ws = cb.Workspace(B, dev)
    ref = None
    for admm in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("f64", "f32")):
        for grace in graces:
            out = None
            for _ in range(2):
                out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=grace, admm_dtype=admm)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=grace, admm_dtype=admm)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 5
            it = out.iters.cpu().numpy()
            st = (out.status & 0xFF).cpu().numpy()
            nstrag = int(ws.buf[:4].view(torch.int64)[1].item())
            R = out.R[:, 0].cpu().numpy()
            if ref is None:
                ref = (R, st)
            both = (st == 0) & (ref[1] == 0)
            ang = synth.rotation_angle(ref[0][both], R[both]).max()
            print(f"{npt}+{nl} admm {admm} grace {grace}: {ms:.2f} ms  iters mean {it.mean():.1f} max {it.max()} "
                  f"status {np.bincount(st, minlength=3)[:3]} handed {nstrag} max rot diff vs first {ang:.1e}", flush=True)
