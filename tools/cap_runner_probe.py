"""Dev tool: where in the work queue do the problems of a batch that run into the iteration cap sit?
Reads the queue order out of the workspace (layout of solve_impl, pnpl_kernels.cu)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B = 100000
dev = torch.device("cuda", 0)
d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=42)
K = torch.from_numpy(d["K"]).to(dev)
args = {}
if n_pts:
    args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
if n_lines:
    args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
ws = cb.Workspace(B, dev)
out = cb.solve_batched(K, **args, workspace=ws, admm_dtype=os.environ.get("ADMM", "f64"))
torch.cuda.synchronize()
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
slots = n_sm * 128
off = (16 + 64) + slots * 45 + B * (112 + 156) + slots * 216          # doubles up to the Anderson words
off_bytes = off * 8 + slots * 512 * 4 + B * 166 * 8 + slots * 2 * 45 * 4
order = ws.buf.view(torch.uint8)[off_bytes:off_bytes + 4 * B].view(torch.int32).cpu().numpy()
assert sorted(order.tolist()) == list(range(B)), "layout drifted"
pos = np.empty(B, np.int64); pos[order] = np.arange(B)
it = out.iters.cpu().numpy(); st = (out.status & 0xFF).cpu().numpy()
slow = np.argsort(-it)[:12]
print("slowest problems: index, iterations, status, queue position (of %d)" % B)
for b in slow:
    print(int(b), int(it[b]), int(st[b]), int(pos[b]))
print("mean queue position of the 1000 slowest:", pos[np.argsort(-it)[:1000]].mean(), " of all:", pos.mean())
