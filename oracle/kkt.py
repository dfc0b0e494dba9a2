"""oracle/kkt.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Solver-independent optimality certificate for the reference's SDP
(cvxpnpl.py:387-448, 475-489):   min <Q, Z>  s.t.  <P_k, Z> = b_k (22 rows), Z >= 0.

Given Q and a candidate Z it reports
  * primal infeasibility  max_k |<P_k, Z> - b_k|  and  max(0, -lambda_min(Z));
  * for a supplied multiplier vector y (22 equalities, SCS sign convention):
    dual infeasibility  max(0, -lambda_min(Q + sum_k y_k P_k)),  complementarity
    ||S Z||  and the gap  <Q, Z> - (-b'y).
Small values of all of them prove Z optimal whatever produced it (SCS, the
restated SCS in scs_port.c, or the CUDA ADMM kernel).
"""
import numpy as np

from .cvxpnpl_oracle import _A, _b, vech_index


def _P():
    Ps = np.zeros((22, 10, 10))
    for k in range(22):
        for j in range(10):
            for i in range(j, 10):
                c = _A[k, vech_index(i, j)]
                if i == j:
                    Ps[k, i, i] = c
                else:
                    Ps[k, i, j] = Ps[k, j, i] = 0.5 * c
    return Ps


P = _P()
b = _b[:22].copy()


def certificate(Q, Z, y=None):
    """y: optional multipliers of the 22 equalities in SCS's sign convention
    (A'y + c = 0, dual objective -b'y), e.g. scs_port's y[:22].  The dual slack
    is then S = Q + sum_k y_k P_k.  Without y a least-squares guess from
    complementary slackness is used (only meaningful when Z is strictly
    complementary; it is NOT a proof of sub-optimality when it fails)."""
    Q = np.asarray(Q, float)
    Z = np.asarray(Z, float)
    eq = np.einsum("kij,ij->k", P, Z) - b
    lam_min = np.linalg.eigvalsh(0.5 * (Z + Z.T))[0]
    if y is None:
        G = np.stack([(P[k] @ Z).ravel() for k in range(22)], axis=1)
        y, *_ = np.linalg.lstsq(G, -(Q @ Z).ravel(), rcond=None)
    y = np.asarray(y, float)[:22]
    S = Q + np.einsum("k,kij->ij", y, P)
    s_min = np.linalg.eigvalsh(0.5 * (S + S.T))[0]
    pobj = float(np.sum(Q * Z))
    dobj = float(-(y @ b))
    return {
        "eq_res": float(np.max(np.abs(eq))),
        "psd_res": float(max(0.0, -lam_min)),
        "dual_psd_res": float(max(0.0, -s_min)),
        "comp_res": float(np.linalg.norm(S @ Z)),
        "pobj": pobj,
        "dobj": dobj,
        "gap": abs(pobj - dobj),
    }
