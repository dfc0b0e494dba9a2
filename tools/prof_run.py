"""Profiling driver: N calls of the batched solve on one synthetic batch (for ncu)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n-pts", type=int, default=8)
ap.add_argument("--n-lines", type=int, default=4)
ap.add_argument("--batch", type=int, default=100000)
ap.add_argument("--calls", type=int, default=2)
ap.add_argument("--handoff", type=int, default=0)
ap.add_argument("--admm", default="f64")
ap.add_argument("--noise", type=float, default=1.0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
d = synth.make_batch(a.batch, a.n_pts, a.n_lines, noise=a.noise, seed=42)
K = torch.from_numpy(d["K"]).to(dev)
args = {}
if a.n_pts:
    args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
if a.n_lines:
    args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
ws = cb.Workspace(a.batch, dev)
out = None
for _ in range(a.calls):
    out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=a.handoff, admm_dtype=a.admm)
torch.cuda.synchronize()
print("iters mean", float(out.iters.float().mean()), "max", int(out.iters.max()))
