"""oracle/candidate_sets.py -- TEST INFRASTRUCTURE (checker), not product code.

Order-free comparison of candidate-pose sets with the reference's (tests/golden/degenerate_big.npz).

The rank-4 recovery (cvxpnpl.py:156-218) is ill conditioned at near-double roots of its
resultant quartic: the reference itself does not reproduce those candidates when Z changes in
the 14th digit.  `stable_mask` marks the reference candidates that it DOES reproduce (they move
less than STABLE_TOL under each of the stored 1e-14 perturbations); parity is asserted on
those, the rest is only counted."""
import numpy as np

STABLE_TOL = 1e-7


def flat(R, t, n):
    return np.concatenate([R[:n].reshape(n, 9), t[:n]], axis=1)


def stable_mask(exp, Rp, tp, n_p):
    """exp [n,12]; Rp [P,4,3,3], tp [P,4,3], n_p [P] -> bool [n]"""
    ok = np.ones(len(exp), bool)
    for k in range(len(n_p)):
        if n_p[k] != len(exp):
            return np.zeros(len(exp), bool)
        per = flat(Rp[k], tp[k], int(n_p[k]))
        ok &= np.abs(exp[:, None, :] - per[None, :, :]).max(-1).min(1) < STABLE_TOL
    return ok


def compare(got, exp, stable):
    """max distance of the stable reference candidates to their nearest returned candidate"""
    if not stable.any():
        return 0.0
    d = np.abs(exp[stable][:, None, :] - got[None, :, :]).max(-1).min(1)
    return float(d.max())
