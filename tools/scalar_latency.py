"""Dev tool: wall-clock latency of the scalar drop-in API (one problem per call)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cvxpnpl_b200 as cb
from cvxpnpl_b200 import synth
d = synth.make_batch(64, 8, 4, noise=1.0, seed=3)
for _ in range(5):
    cb.pnpl(d["pts_2d"][0], d["line_2d"][0], d["pts_3d"][0], d["line_3d"][0], d["K"])
t0 = time.perf_counter()
for i in range(64):
    cb.pnpl(d["pts_2d"][i], d["line_2d"][i], d["pts_3d"][i], d["line_3d"][i], d["K"])
print(f"scalar pnpl: {(time.perf_counter() - t0) / 64 * 1e3:.3f} ms per call")
