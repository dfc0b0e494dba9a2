"""Dev tool: tracked-eigenpair solver (psd="track") against the full-decomposition solver (psd="full") on the
same seeded batches: step time, per-kernel times, iteration counts, statuses, pose agreement."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth

dev = torch.device("cuda", 0)
configs = [(8, 4, 100000, "f64"), (8, 0, 100000, "f64"), (0, 6, 100000, "f64"), (0, 6, 100000, "f32"), (4, 0, 20000, "f64"),
           (6, 0, 50000, "f64")]
if len(sys.argv) > 1:
    configs = [c for c in configs if f"{c[0]}+{c[1]}" in sys.argv[1:]]
out_rows = []
for n_pts, n_lines, B, admm in configs:
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=42)
    K = torch.from_numpy(d["K"]).to(dev)
    args = {}
    if n_pts:
        args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
    if n_lines:
        args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    ws = cb.Workspace(B, dev)
    res = {}
    row = dict(n_pts=n_pts, n_lines=n_lines, B=B, admm=admm)
    for psd in ("full", "track"):
        out = None
        for _ in range(2):
            out = cb.solve_batched(K, **args, workspace=ws, out=out, psd=psd, admm_dtype=admm)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            out = cb.solve_batched(K, **args, workspace=ws, out=out, psd=psd, admm_dtype=admm, timing=True)
        e.record()
        torch.cuda.synchronize()
        t = {k: round(v, 3) for k, v in cb.last_kernel_times().items() if v > 0}
        it = out.iters.cpu().numpy()
        st = (out.status & 0xFF).cpu().numpy()
        res[psd] = dict(R=out.R.cpu().numpy().copy(), t=out.t.cpu().numpy().copy(), n=out.n_poses.cpu().numpy().copy(), st=st,
                        it=it)
        nfail = int(ws.buf[:16].view(torch.int64)[6].item())
        row[psd] = dict(ms=round(s.elapsed_time(e) / 5, 3), kernels=t, iters_mean=float(it.mean()), iters_max=int(it.max()),
                        status_hist=np.bincount(st, minlength=5).tolist(), handed_back=nfail, launches=out.launches)
    a, b = res["full"], res["track"]
    ok = (a["st"] == 0) & (b["st"] == 0) & (a["n"] == 1) & (b["n"] == 1)
    ang = synth.rotation_angle(a["R"][ok, 0], b["R"][ok, 0])
    terr = np.linalg.norm(a["t"][ok, 0] - b["t"][ok, 0], axis=1) / np.linalg.norm(a["t"][ok, 0], axis=1)
    row["agree"] = dict(both_ok_frac=float(ok.mean()), status_diff_frac=float((a["st"] != b["st"]).mean()),
                        n_poses_diff_frac=float((a["n"] != b["n"]).mean()), max_rot=float(ang.max()), max_t=float(terr.max()))
    print(json.dumps(row), flush=True)
    out_rows.append(row)
