"""BASELINE.json configs[4]: degenerate / ambiguous sweep, 1e4 problems on one GPU.

Six noise-free families of ~1667 problems each: 3 points,
4 points, 3 lines, 2 points + 1 line, 4 lines, 8 coplanar points.  Reports, per
family, the status and candidate-count histograms, iterations, the fraction of
converged problems whose candidate set contains the ground truth, and the device time
of the batched solve (CUDA events).  One JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth

FAMILIES = [("pts3", 3, 0, False), ("pts4", 4, 0, False), ("lines3", 0, 3, False), ("p2l1", 2, 1, False),
            ("lines4", 0, 4, False), ("coplanar8", 8, 0, True)]
dev = torch.device("cuda", 0)
out, total_ms, total_n = {}, 0.0, 0
for name, n_pts, n_lines, coplanar in FAMILIES:
    B = 1667 if name != "coplanar8" else 1665
    d = synth.make_batch(B, n_pts, n_lines, noise=0.0, seed=9, coplanar=coplanar)
    K = torch.from_numpy(d["K"]).to(dev)
    args = {}
    if n_pts:
        args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
    if n_lines:
        args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    ws = cb.Workspace(B, dev)
    res = None
    for _ in range(3):
        res = cb.solve_batched(K, **args, workspace=ws, out=res)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        res = cb.solve_batched(K, **args, workspace=ws, out=res)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    st = (res.status & 0xFF).cpu().numpy()
    n = res.n_poses.cpu().numpy()
    it = res.iters.cpu().numpy()
    R = res.R.cpu().numpy()
    fin = np.isfinite(R).all(axis=(2, 3))
    ang = np.where(fin, synth.rotation_angle(d["R_gt"][:, None], R), np.inf)
    conv = st == 0
    out[name] = {"problems": B, "ms": ms, "status_hist[ok,max_iters,nan,singular,rank0]": np.bincount(st, minlength=5).tolist(),
                 "n_poses_hist[0,1,2,3,4]": np.bincount(n, minlength=5).tolist(),
                 "iters_median": float(np.median(it)), "iters_max": int(it.max()),
                 "gt_among_candidates_of_converged": float((ang.min(axis=1) < 1e-4)[conv].mean()) if conv.any() else None}
    total_ms += ms
    total_n += B
print(json.dumps({"workload": "BASELINE.json configs[4]: degenerate sweep, noise free, fp64", "problems": total_n,
                  "ms_total": total_ms, "poses_per_s": total_n / (total_ms * 1e-3), "families": out}))
