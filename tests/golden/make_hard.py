"""Generator of tests/golden/hard_pnpl.npz: the slowest PnPL (8 points + 4 lines) problems of three seeded
1e5-problem synthetic batches (seeds 47, 52, 42; sigma = 1 px), found with the host build of the solver
routines (tests/host) with the plateau jump switched off.  They are the problems whose DR iteration
walks along a plateau for hundreds of iterations (799, 738, 469, ... iterations without the jump); the
tests pin the iteration counts with the jump and the poses against the oracle.

    python tests/golden/make_hard.py        # ~2 min on 8 cores
"""
import os
import subprocess
import sys
import ctypes
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvxpnpl_b200 import synth  # noqa: E402

KEYS = ("pts_2d", "pts_3d", "line_2d", "line_3d", "R_gt", "t_gt")
SO = "/tmp/libhost_noplat.so"


def _lib():
    if not os.path.exists(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DCVX_PLAT_N=100000", "-o", SO,
                               os.path.join(ROOT, "tests", "host", "host_harness.cpp")])
    from tests.host import harness
    harness._lib = ctypes.CDLL(SO)
    return harness


def work(a):
    seed, lo, hi = a
    d = synth.make_batch(100000, 8, 4, noise=1.0, seed=seed)
    sub = {k: (d[k][lo:hi] if k in KEYS else d[k]) for k in d}
    return lo, _lib().solve(sub)["iters"]


if __name__ == "__main__":
    _lib()
    out = {}
    for seed in (47, 52, 42):
        with Pool(8) as p:
            res = p.map(work, [(seed, lo, lo + 12500) for lo in range(0, 100000, 12500)])
        it = np.concatenate([r[1] for r in sorted(res, key=lambda r: r[0])])
        d = synth.make_batch(100000, 8, 4, noise=1.0, seed=seed)
        hard = np.argsort(-it)[:4]
        print(seed, "slowest:", it[hard])
        for k in KEYS:
            out.setdefault(k, []).append(d[k][hard])
        out.setdefault("iters_without_jump", []).append(it[hard])
        out["K"] = d["K"]
    np.savez(os.path.join(ROOT, "tests", "golden", "hard_pnpl.npz"),
             **{k: (np.concatenate(v) if isinstance(v, list) else v) for k, v in out.items()})
