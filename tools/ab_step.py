"""Dev tool: headline step time + per-kernel times for the library named by CVXPNPL_B200_LIB (A/B builds).
usage: ab_step.py [n_pts n_lines [batch [psd]]]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
psd = sys.argv[4] if len(sys.argv) > 4 else "track"
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 42
admm = os.environ.get("ADMM", "f64")
noise = float(os.environ.get("NOISE", "1.0"))
dev = torch.device("cuda", 0)
d = synth.make_batch(B, n_pts, n_lines, noise=noise, seed=seed)
K = torch.from_numpy(d["K"]).to(dev)
args = {}
if n_pts:
    args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
if n_lines:
    args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
ws = cb.Workspace(B, dev)
out = None
for _ in range(3):
    out = cb.solve_batched(K, **args, workspace=ws, out=out, psd=psd, admm_dtype=admm)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(8):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = cb.solve_batched(K, **args, workspace=ws, out=out, psd=psd, timing=True, admm_dtype=admm)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
t = {k: round(v, 3) for k, v in cb.last_kernel_times().items() if v > 0}
it = out.iters.cpu().numpy()
st = (out.status & 0xFF).cpu().numpy()
print(f"{os.environ.get('CVXPNPL_B200_LIB', 'default'):28s} {n_pts}+{n_lines} B={B} {psd}: {np.median(ts):.3f} ms (min {min(ts):.3f})  {t}  "
      f"iters mean {it.mean():.2f} max {it.max()} status {np.bincount(st, minlength=5).tolist()} handed_back {int(ws.buf[:16].view(torch.int64)[6].item())}", flush=True)
