"""Generates tests/golden/degenerate_big.npz: a larger replay set for the multi-solution
extraction (cvxpnpl.py:221-343, 156-218) and its error paths (LinAlgError at cvxpnpl.py:165,
212 and from np.linalg.svd at 510 when a candidate is not finite).  Run from the repo root:

    python tests/golden/make_degenerate_big.py

The VERBATIM reference (/root/reference/cvxpnpl.py) runs with oracle/shim/scs standing in
for SCS at the reference's own defaults (eps 1e-9, max_iters 2500); what is stored per
problem is the Z the reference saw, its A'A / B, and what the reference then did with that
Z: the candidate poses, or the exception it raised.  Solver independent by construction.

The rank-4 recovery is ill conditioned where the resultant quartic (cvxpnpl.py:176-185) has a
near-double root: those candidates move by 1e-6 ... 1e-1 when Z changes in the 14th digit, i.e.
the reference does not reproduce them itself across LAPACK builds.  To make "parity" well
defined the reference is therefore ALSO run on N_PERT copies of Z perturbed by 1e-14 (relative,
symmetric Gaussian); the tests assert parity on every candidate the reference reproduces under
that perturbation (moves less than 1e-7) and only report the others.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cvxpnpl as ref  # noqa: E402  (the verbatim reference)
import scs  # noqa: E402  (the shim)
from cvxpnpl_b200 import synth  # noqa: E402
from make_golden import _call, _ref_AB, pack_poses  # noqa: E402

warnings.simplefilter("ignore")
N_PER_FAMILY = 40
N_PERT = 3
PERT = 1e-14
FAMILIES = {"pts4": (4, 0, False), "pts3": (3, 0, False), "lines3": (0, 3, False), "lines4": (0, 4, False),
            "p2l1": (2, 1, False), "coplanar8": (8, 0, True), "pts5": (5, 0, False)}


def main():
    captured = {}
    real_solve = scs.solve

    def spy(data, cone, **kw):
        if "replay" in captured:          # second pass: hand the reference a perturbed copy of the same x
            return {"x": captured["replay"], "info": {"dobj": 0.0}}
        res = real_solve(data, cone, **kw)
        captured["x"] = res["x"].copy()
        return res

    rng = np.random.default_rng(5)

    scs.solve = spy
    ref.scs.solve = spy
    out = {}
    for name, (n_pts, n_lines, cop) in FAMILIES.items():
        d = synth.make_batch(N_PER_FAMILY, n_pts, n_lines, noise=0.0, seed=78, coplanar=cop)
        Rs, ts, ns, Zs, As, Bs, err = [], [], [], [], [], [], []
        Rp, tp, npert = [], [], []
        for i in range(N_PER_FAMILY):
            A, Bm = _ref_AB(d, i, n_pts, n_lines)
            try:
                R, t, n = pack_poses(_call(d, i, n_pts, n_lines))
                err.append(0)
            except np.linalg.LinAlgError:
                R, t, n = np.full((4, 3, 3), np.nan), np.full((4, 3), np.nan), 0
                err.append(1)
            except NotImplementedError:
                R, t, n = np.full((4, 3, 3), np.nan), np.full((4, 3), np.nan), 0
                err.append(2)
            Rs.append(R), ts.append(t), ns.append(n)
            Zs.append(ref._vech10_inv(captured["x"]))
            As.append(A.T @ A), Bs.append(Bm)
            x0 = captured["x"]
            Rk, tk, nk = [], [], []
            for _ in range(N_PERT):
                captured["replay"] = x0 * (1.0 + PERT * rng.standard_normal(x0.shape))
                try:
                    R, t, n = pack_poses(_call(d, i, n_pts, n_lines))
                except (np.linalg.LinAlgError, NotImplementedError):
                    R, t, n = np.full((4, 3, 3), np.nan), np.full((4, 3), np.nan), 0
                Rk.append(R), tk.append(t), nk.append(n)
            del captured["replay"]
            Rp.append(Rk), tp.append(tk), npert.append(nk)
        out[f"{name}_R"], out[f"{name}_t"], out[f"{name}_n"] = np.array(Rs), np.array(ts), np.array(ns, dtype=np.int32)
        out[f"{name}_Z"], out[f"{name}_AtA"], out[f"{name}_B"] = np.array(Zs), np.array(As), np.array(Bs)
        out[f"{name}_err"] = np.array(err, dtype=np.int32)   # 0 poses returned, 1 LinAlgError, 2 NotImplementedError
        out[f"{name}_Rp"], out[f"{name}_tp"] = np.array(Rp), np.array(tp)      # [N, N_PERT, 4, 3, 3], [N, N_PERT, 4, 3]
        out[f"{name}_np"] = np.array(npert, dtype=np.int32)
        print(name, "n_poses", np.bincount(ns, minlength=5).tolist(), "errors", np.bincount(err, minlength=3).tolist())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "degenerate_big.npz"), **out)


if __name__ == "__main__":
    main()
