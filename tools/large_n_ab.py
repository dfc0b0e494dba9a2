"""Dev tool: large-n assembly, TMA-staged kernel vs plain-load kernel: agreement and achieved bandwidth."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import suite
dev = torch.device("cuda", 0)
for n, B in ((10000, 4000), (2000, 20000), (1000000, 40)):
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    d = suite.generate(B, n, 0, 1.0, gen, dev)
    p2, p3, K = d["pts_2d"], d["pts_3d"], d["K"]
    out = {}
    for staging in ("loads", "tma"):
        for _ in range(3):
            Q, Bm = cb.assemble_batched(K, p2, p3, staging=staging)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); Q, Bm = cb.assemble_batched(K, p2, p3, staging=staging); e.record()
            torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        out[staging] = (Q.clone(), Bm.clone(), min(ts))
    dq = (out["tma"][0] - out["loads"][0]).abs().max().item() / out["loads"][0].abs().max().item()
    db = (out["tma"][1] - out["loads"][1]).abs().max().item() / out["loads"][1].abs().max().item()
    gb = 40.0 * n * B / 1e9
    print(f"n={n} B={B}: loads {out['loads'][2]:.3f} ms ({gb / out['loads'][2] * 1e3:.0f} GB/s)  tma {out['tma'][2]:.3f} ms ({gb / out['tma'][2] * 1e3:.0f} GB/s)  rel diff Q {dq:.1e} B {db:.1e}", flush=True)
