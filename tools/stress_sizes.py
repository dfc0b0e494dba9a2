"""Dev tool: odd batch sizes / configurations through every path; default options vs the
plain thread path (handoff=-1) vs the FP32 first phase must agree on converged problems."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
dev = torch.device("cuda", 0)
bad = 0
for (npt, nl) in ((8, 4), (0, 6), (5, 0), (3, 2), (12, 7)):
    for B in (1, 2, 31, 127, 128, 129, 2367, 2368, 2369, 18943, 18944, 18945, 40001):
        d = synth.make_batch(B, npt, nl, noise=1.0, seed=B + npt)
        kb = (B % 2 == 1)
        K = np.broadcast_to(d["K"], (B, 3, 3)).copy() if kb else d["K"]
        args = {}
        if npt: args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
        if nl: args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
        Kd = torch.from_numpy(np.ascontiguousarray(K)).to(dev)
        ref = cb.solve_batched(Kd, **args, handoff=-1)
        for mode in ({}, {"admm_dtype": "f32"}, {"handoff": 3}, {"variant": "rc"}):
            r = cb.solve_batched(Kd, **args, **mode)
            torch.cuda.synchronize()
            st = (r.status & 0xFF).cpu().numpy()
            if "variant" in mode:
                ok = np.isin(st, (0, 1, 3, 4)).all()
                if not ok: bad += 1; print("BAD status", npt, nl, B, mode)
                continue
            both = (st == 0) & ((ref.status & 0xFF).cpu().numpy() == 0) & (r.n_poses.cpu().numpy() == 1) & (ref.n_poses.cpu().numpy() == 1)
            if both.any():
                ang = synth.rotation_angle(ref.R[:, 0].cpu().numpy()[both], r.R[:, 0].cpu().numpy()[both]).max()
                if not (ang < 1e-6): bad += 1; print("BAD pose", npt, nl, B, mode, ang)
            frac = both.mean()
            if frac < (0.5 if npt + nl < 8 else 0.9): print("low agreement fraction", npt, nl, B, mode, frac)
    print("done", npt, nl, flush=True)
print("bad =", bad)
