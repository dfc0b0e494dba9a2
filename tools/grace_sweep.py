import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
dev = torch.device('cuda', 0)
for (npt, nl) in ((8, 4), (8, 0), (0, 6)):
    B = 100000
    d = synth.make_batch(B, npt, nl, noise=1.0, seed=42)
    K = torch.from_numpy(d["K"]).to(dev)
    args = {}
    if npt: args.update(pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev))
    if nl: args.update(line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
    ws = cb.Workspace(B, dev)
    for grace in (-1, 1, 10, 20, 40, 80, 160):
        out = None
        for _ in range(2): out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=grace)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): out = cb.solve_batched(K, **args, workspace=ws, out=out, handoff=grace)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        it = out.iters.cpu().numpy(); st = (out.status & 0xff).cpu().numpy()
        nstrag = int(ws.buf[:4].view(torch.int64)[1].item())
        print(f"{npt}+{nl} grace {grace}: {ms:.2f} ms  iters mean {it.mean():.1f} max {it.max()} status {np.bincount(st, minlength=3)[:3]} handed {nstrag}", flush=True)
