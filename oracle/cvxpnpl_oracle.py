"""oracle/cvxpnpl_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy, fp64) restatement of the reference's pnp / pnl / pnpl path
(/root/reference/cvxpnpl.py, commit e20cca87).  It exists because the reference
cannot travel to the GPU box; it is the checker the CUDA path is compared with.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import it.  The product package (cvxpnpl_b200/) never does.

Pinning:
  * every function here is checked against the VERBATIM reference functions
    (imported from /root/reference in the build container) on seeded inputs, and
    the resulting input/output vectors are committed under tests/golden/ by
    tests/golden/make_golden.py; tests/test_oracle.py replays them without the
    reference;
  * the three known-answer examples (examples/pnp.py, pnl.py, pnpl.py);
  * the SDP solve itself (`scs.solve`, cvxpnpl.py:485-489) is third-party SCS,
    absent from the reference tree and from this image: PARITY UNPINNED at that
    boundary.  oracle/scs_port.c restates the published SCS algorithm and
    oracle/kkt.py certifies any candidate optimum independently of any solver.

Each function cites the reference lines it follows.  Formulas are restated, not
copied: the constraint builders use the Kronecker structure of the rows, the
static SDP data is generated from the constraint table in SURVEY.md section 3.2,
the 21 quadratic forms are built with einsum, and the E6Q3 quartic is obtained
as det M(a) by polynomial arithmetic instead of the reference's expanded
640-term coefficient expressions (same polynomial up to sign, see _e6q3).
"""
import warnings

import numpy as np

from . import scs_port

# ----------------------------------------------------------------------------
# static SDP data (cvxpnpl.py:387-451)
# ----------------------------------------------------------------------------

#: the 15 "triple" equalities, each  sum_k sign_k * Z[i_k, j_k] = 0  over
#: off-diagonal entries.  Order = rows 2,3,5 (row orthogonality), 8,9,11 (column
#: orthogonality), 13..21 (c1 x c2 = c3, c2 x c3 = c1, c3 x c1 = c2) of the
#: reference's A (cvxpnpl.py:401-435); r index i <-> R[i % 3, i // 3].
TRIPLES = {
    2: ((0, 1, 1), (3, 4, 1), (6, 7, 1)),
    3: ((0, 2, 1), (3, 5, 1), (6, 8, 1)),
    5: ((1, 2, 1), (4, 5, 1), (7, 8, 1)),
    8: ((0, 3, 1), (1, 4, 1), (2, 5, 1)),
    9: ((0, 6, 1), (1, 7, 1), (2, 8, 1)),
    11: ((3, 6, 1), (4, 7, 1), (5, 8, 1)),
    13: ((1, 5, 1), (2, 4, -1), (6, 9, -1)),
    14: ((2, 3, 1), (0, 5, -1), (7, 9, -1)),
    15: ((0, 4, 1), (1, 3, -1), (8, 9, -1)),
    16: ((4, 8, 1), (5, 7, -1), (0, 9, -1)),
    17: ((5, 6, 1), (3, 8, -1), (1, 9, -1)),
    18: ((3, 7, 1), (4, 6, -1), (2, 9, -1)),
    19: ((2, 7, 1), (1, 8, -1), (3, 9, -1)),
    20: ((0, 8, 1), (2, 6, -1), (4, 9, -1)),
    21: ((1, 6, 1), (0, 7, -1), (5, 9, -1)),
}
#: the 6 norm equalities  Z[i,i] + Z[j,j] + Z[k,k] - Z[9,9] = 0
NORMS = {1: (0, 3, 6), 4: (1, 4, 7), 6: (2, 5, 8), 7: (0, 1, 2), 10: (3, 4, 5), 12: (6, 7, 8)}


def vech_index(i, j):
    """Position of Z[i, j] in the column-major lower-triangle vectorisation
    (cvxpnpl.py:356-369)."""
    if i < j:
        i, j = j, i
    return 10 * j - j * (j - 1) // 2 + (i - j)


def vech10(S, scale=1.0):
    """cvxpnpl.py:346-370: symmetric 10x10 -> 55, off-diagonals times `scale`."""
    out = np.empty(55)
    for j in range(10):
        for i in range(j, 10):
            out[vech_index(i, j)] = S[i, j] * (1.0 if i == j else scale)
    return out


def vech10_inv(v):
    """cvxpnpl.py:373-384: 55 -> symmetric 10x10 (no unscaling)."""
    Z = np.empty((10, 10))
    for j in range(10):
        for i in range(j, 10):
            Z[i, j] = Z[j, i] = v[vech_index(i, j)]
    return Z


def sdp_constraints():
    """cvxpnpl.py:387-448: dense A (77x55) and b (77) of
    min c'x s.t. Ax + s = b, s in {0}^22 x S_+^10.  Each off-diagonal Z_ij
    enters an equality row with coefficient +-1 (P symmetrised, vech scale 2);
    the cone block is -diag(1 | sqrt 2)."""
    A = np.zeros((77, 55))
    A[0, vech_index(9, 9)] = 1.0
    for row, idx in NORMS.items():
        for i in idx:
            A[row, vech_index(i, i)] = 1.0
        A[row, vech_index(9, 9)] = -1.0
    for row, tri in TRIPLES.items():
        for i, j, s in tri:
            A[row, vech_index(i, j)] = float(s)
    for j in range(10):
        for i in range(j, 10):
            k = vech_index(i, j)
            A[22 + k, k] = -1.0 if i == j else -np.sqrt(2.0)
    b = np.zeros(77)
    b[0] = 1.0
    return A, b


_A, _b = sdp_constraints()


def sdp_constraints_rc():
    """benchmarks/toolkit/methods/rc.py:9-60: the "rc" ablation drops the six row
    orthonormality equalities (rows 1-6 of the full A: norms 1,4,6 and triples
    2,3,5), leaving 16 equalities: A is 71 x 55."""
    drop = [1, 2, 3, 4, 5, 6]
    keep = [k for k in range(77) if k not in drop]
    return _A[keep].copy(), _b[keep].copy()


_A_rc, _b_rc = sdp_constraints_rc()

# ----------------------------------------------------------------------------
# constraint builders (cvxpnpl.py:20-153)
# ----------------------------------------------------------------------------


def _skew(p):
    """[p]_x for rows of p: (n,3) -> (n,3,3)."""
    n = len(p)
    S = np.zeros((n, 3, 3))
    S[:, 0, 1], S[:, 0, 2] = -p[:, 2], p[:, 1]
    S[:, 1, 0], S[:, 1, 2] = p[:, 2], -p[:, 0]
    S[:, 2, 0], S[:, 2, 1] = -p[:, 1], p[:, 0]
    return S


def point_constraints(pts_2d, pts_3d, K):
    """cvxpnpl.py:20-104.  Bearing p = K^-1 [u v 1]' (line 37); the three
    constraint rows of point i are  P_i' (x) row_k([p_i]_x)  (lines 53-98) and
    the translation rows are row_k([p_i]_x) (lines 83-102).  Returns the
    reference's stacking: C (3n, 9) = [C1; C2; C3], N (3n, 3) = [N1; N2; N3]."""
    pts_2d = np.asarray(pts_2d, float).reshape(-1, 2)
    pts_3d = np.asarray(pts_3d, float).reshape(-1, 3)
    n = len(pts_3d)
    hom = np.vstack((pts_2d.T, np.ones(n)))
    p = np.linalg.solve(np.asarray(K, float), hom).T  # (n, 3)
    S = _skew(p)  # (n, k, :)
    # C_k[i] = kron(P_i, S[i, k, :])
    C = np.einsum("na,nkb->knab", pts_3d, S).reshape(3, n, 9)
    N = S.transpose(1, 0, 2)  # (k, n, 3)
    return C.reshape(3 * n, 9), N.reshape(3 * n, 3)


def line_constraints(line_2d, line_3d, K):
    """cvxpnpl.py:107-153.  Normal of the back-projected plane
    n = normalise(K^-1 p0 x K^-1 p1) (lines 123-132), one row per 3D endpoint:
    C row = kron(P, n), N row = n (lines 133-153)."""
    line_2d = np.asarray(line_2d, float).reshape(-1, 2, 2)
    line_3d = np.asarray(line_3d, float).reshape(-1, 2, 3)
    m = len(line_2d)
    hom = np.vstack((line_2d.reshape(2 * m, 2).T, np.ones(2 * m)))
    l = np.linalg.solve(np.asarray(K, float), hom).T.reshape(m, 2, 3)
    nrm = np.cross(l[:, 0], l[:, 1])
    nrm = nrm / np.linalg.norm(nrm, axis=1)[:, None]
    nn = np.repeat(nrm, 2, axis=0)  # (2m, 3), one per endpoint
    P = line_3d.reshape(2 * m, 3)
    C = np.einsum("na,nb->nab", P, nn).reshape(2 * m, 9)
    return C, nn


def reduce_translation(C, N):
    """cvxpnpl.py:548-549 / 579-580 / 623-624: B = (N'N)^-1 N'C, A = C - N B."""
    B = np.linalg.solve(N.T @ N, N.T @ C)
    A = C - N @ B
    return A, B


# ----------------------------------------------------------------------------
# multi-solution extraction (cvxpnpl.py:156-343)
# ----------------------------------------------------------------------------


def _polymul(p, q):
    return np.convolve(p, q)


def _e6q3(A):
    """cvxpnpl.py:156-218.  A is (N,10) over monomials
    [a^2, b^2, c^2, ab, ac, bc, a, b, c, 1].

    Lines 163-173: least-squares express the 6 quadratic monomials through
    (a, b, c, 1); keep  b^2 = d0.(a,b,c,1), c^2 = d1.(a,b,c,1), bc = d2.(a,b,c,1).
    Lines 176-181 then hard-code the expanded resultant quartic in a.  Here it is
    restated as the determinant of the hidden-variable matrix M(a) obtained from
    the identities  b(bc) = c(b^2), c(bc) = b(c^2), (bc)^2 = b^2 c^2  reduced once
    more through the three relations (derived symbolically; M equals minus the
    reference's M of lines 190-202, so det M(a) is minus the reference quartic and
    has the same roots).  Lines 185-186: all four roots, real part taken; lines
    205-216: (b, c) by least squares on  M[:, :2] (b, c)' = -M[:, 2].
    """
    Bq, Cl = A[:, :6], A[:, 6:]
    X = np.linalg.solve(Bq.T @ Bq, Bq.T @ Cl)
    D = -X[[1, 2, 5]]
    (d00, d01, d02, d03), (d10, d11, d12, d13), (d20, d21, d22, d23) = D

    # M(a) = M0 + a M1 + a^2 M2, rows = identities, columns = (b, c, 1)
    M0 = np.array([
        [-d02 * d11 + d21 * d22 + d23,
         -d01 * d22 - d02 * d12 + d02 * d21 - d03 + d22 ** 2,
         -d01 * d23 - d02 * d13 + d03 * d21 + d22 * d23],
        [-d01 * d11 + d11 * d22 - d12 * d21 - d13 + d21 ** 2,
         -d02 * d11 + d21 * d22 + d23,
         -d03 * d11 - d12 * d23 + d13 * d22 + d21 * d23],
        [-d01 ** 2 * d11 - d01 * d12 * d21 - d01 * d13 + d01 * d21 ** 2 - d02 * d11 * d12
         - d02 * d11 * d21 - d03 * d11 + d11 * d22 ** 2 + 2 * d21 ** 2 * d22 + 2 * d21 * d23,
         -d01 * d02 * d11 - d01 * d12 * d22 - d02 * d11 * d22 - d02 * d12 ** 2 - d02 * d13
         + d02 * d21 ** 2 - d03 * d12 + d12 * d22 ** 2 + 2 * d21 * d22 ** 2 + 2 * d22 * d23,
         -d01 * d03 * d11 - d01 * d12 * d23 - d02 * d11 * d23 - d02 * d12 * d13 - d03 * d13
         + d03 * d21 ** 2 + d13 * d22 ** 2 + 2 * d21 * d22 * d23 + d23 ** 2],
    ])
    M1 = np.array([
        [d20, -d00, d00 * d21 - d01 * d20 - d02 * d10 + d20 * d22],
        [-d10, d20, -d00 * d11 + d10 * d22 - d12 * d20 + d20 * d21],
        [-d00 * d11 - d01 * d10 + 2 * d20 * d21,
         -d00 * d12 - d02 * d10 + 2 * d20 * d22,
         -d00 * d01 * d11 - d00 * d13 + d00 * d21 ** 2 - d01 * d12 * d20 - d02 * d10 * d12
         - d02 * d11 * d20 - d03 * d10 + d10 * d22 ** 2 + 2 * d20 * d21 * d22 + 2 * d20 * d23],
    ])
    M2 = np.zeros((3, 3))
    M2[2, 2] = -d00 * d10 + d20 ** 2

    # entries as ascending-power polynomials in a
    E = [[np.array([M0[i, j], M1[i, j], M2[i, j]]) for j in range(3)] for i in range(3)]

    def minor(i0, j0, i1, j1):
        return _polymul(E[i0][j0], E[i1][j1]) - _polymul(E[i0][j1], E[i1][j0])

    det = (_polymul(E[0][0], minor(1, 1, 2, 2))
           - _polymul(E[0][1], minor(1, 0, 2, 2))
           + _polymul(E[0][2], minor(1, 0, 2, 1)))  # ascending, degree <= 6
    # structurally degree 4 (only M[2,2] carries a^2 and rows 0,1 are affine in a
    # with a rank-1 a-part in columns (b,c)); drop the vanishing top terms
    quartic = det[:5][::-1]  # descending p4..p0
    a = np.real(np.roots(quartic))

    Ma = M0[None] + a[:, None, None] * M1[None] + (a ** 2)[:, None, None] * M2[None]
    L = Ma[:, :, :2]
    rhs = Ma[:, :, 2, None]
    bc = -np.linalg.solve(L.transpose(0, 2, 1) @ L, L.transpose(0, 2, 1) @ rhs)
    b, c = bc.reshape(-1, 2).T
    return a, b, c


def quadratic_forms(V):
    """cvxpnpl.py:238-301: the 21 symmetric forms P (21, k, k) such that
    alpha' P alpha = 0 (alpha[-1] = 1) encodes, for r = V alpha (column-major vec
    of R): 6 column products c_i.c_j = delta_ij (Pc), 6 row products (Pr), and the
    9 cross-product identities (c_i x c_j)_l = (c_k)_l for (i,j,k) cyclic."""
    k = V.shape[1]
    T = V.reshape(3, 3, k)  # T[col, row, :]
    last = np.zeros((k, k))
    last[-1, -1] = 1.0
    Pc, Pr = [], []
    for i in range(3):
        for j in range(i, 3):
            P = np.einsum("ra,rb->ab", T[i], T[j]) - (i == j) * last
            Pc.append(0.5 * (P + P.T))
            P = np.einsum("ca,cb->ab", T[:, i], T[:, j]) - (i == j) * last
            Pr.append(0.5 * (P + P.T))
    eps = np.zeros((3, 3, 3))
    eps[0, 1, 2] = eps[1, 2, 0] = eps[2, 0, 1] = 1
    eps[0, 2, 1] = eps[2, 1, 0] = eps[1, 0, 2] = -1
    Pd = []
    for i, j, kk in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
        for l in range(3):
            # (c_i x c_j)_l = eps[l,m,n] c_i[m] c_j[n]; the reference writes it as
            # Vcj' S_l Vci with S_l = [e_l]_x  (line 293)
            P = np.einsum("mn,ma,nb->ba", eps[l], T[i], T[j])
            lin = np.zeros((k, k))
            lin[-1, :] = T[kk, l]
            P = P - lin
            Pd.append(0.5 * (P + P.T))
    return np.array(Pc + Pr + Pd)


def constraint_ortho_det(vecs, rank):
    """cvxpnpl.py:221-343."""
    _rank = min(int(np.ceil(rank / 2) * 2), 4)
    Vt = vecs[:, -_rank:].T.copy()  # rows = top eigenvectors (line 234)
    v0 = Vt[-1] / Vt[-1, -1]
    Vt = np.vstack((Vt[:-1] - np.outer(Vt[:-1, -1], v0), v0))
    V = Vt.T[:-1]  # (9, _rank) (lines 235-236)

    P = quadratic_forms(V)
    if _rank == 2:
        # lines 303-315: average the 21 quadratics in a, closed-form roots
        co = np.mean(np.stack([P[:, 0, 0], 2 * P[:, 0, 1], P[:, 1, 1]], axis=1), axis=0)
        root = np.sqrt(max(co[1] * co[1] - 4 * co[0] * co[2], 0.0))
        a = np.array([(-co[1] + root) / (2 * co[0]), (-co[1] - root) / (2 * co[0])])
        alpha = np.stack([a, np.ones(2)], axis=1)
    elif _rank == 4:
        # lines 317-338
        A = np.stack([P[:, 0, 0], P[:, 1, 1], P[:, 2, 2], 2 * P[:, 0, 1], 2 * P[:, 0, 2],
                      2 * P[:, 1, 2], 2 * P[:, 0, 3], 2 * P[:, 1, 3], 2 * P[:, 2, 3],
                      P[:, 3, 3]], axis=1)
        a, b, c = _e6q3(A)
        alpha = np.stack([a, b, c, np.ones(len(a))], axis=1)
    else:
        raise NotImplementedError  # line 341
    return alpha @ V.T


# ----------------------------------------------------------------------------
# relaxation + extraction (cvxpnpl.py:454-520)
# ----------------------------------------------------------------------------


def solve_sdp(Q, eps=1e-9, max_iters=2500, variant="full", eps_rel=0.0):
    """cvxpnpl.py:478-492 (variant "full") or rc.py:86-96 (variant "rc") with the scs
    call replaced by oracle/scs_port.  eps_rel = 0 solves to the absolute tolerance the
    reference asks for; eps_rel = 1e-4 is what real SCS 3.x would add by default (the reference
    leaves it untouched, cvxpnpl.py:17)."""
    A, b = (_A, _b) if variant == "full" else (_A_rc, _b_rc)
    res = scs_port.solve(A, b, vech10(Q, 2.0), eps_abs=eps, eps_rel=eps_rel, max_iters=max_iters)
    info = dict(res["info"])
    info["y"] = res["y"]
    return vech10_inv(res["x"]), info


def extract(Z, A, B, dobj=None, eps=1e-9):
    """cvxpnpl.py:493-520: NaN guard, rank test (lambda > 1e-3), rank-1 or
    multi-solution recovery, SO(3) projection by SVD *without* determinant fix,
    t = -B r, optimality warning."""
    if np.any(np.isnan(Z)):
        return [(np.full((3, 3), np.nan), np.full(3, np.nan))]
    vals, vecs = np.linalg.eigh(Z)
    rank = int(np.sum(vals > 1e-3))
    if rank == 1:
        r_c = (vecs[:-1, -1] / vecs[-1, -1])[None, :]
    else:
        r_c = constraint_ortho_det(vecs, rank)
    U, _, Vh = np.linalg.svd(r_c.reshape(-1, 3, 3))
    Rt = U @ Vh
    r = Rt.reshape(-1, 9)
    t = -r @ B.T
    if dobj is not None:
        res = r @ A.T
        if np.any(np.abs(np.sum(res * res, axis=-1) - dobj) > eps):
            warnings.warn("The solution is not certifiably optimal.")
    return list(zip(Rt.transpose(0, 2, 1), t))


def solve_relaxation(A, B, eps=1e-9, max_iters=2500, return_aux=False, variant="full", eps_rel=0.0):
    """cvxpnpl.py:454-520; variant "rc" = rc.py:65-122 (same extraction, no
    optimality warning)."""
    Q = np.zeros((10, 10))
    Q[:9, :9] = A.T @ A
    Z, info = solve_sdp(Q, eps, max_iters, variant, eps_rel)
    poses = extract(Z, A, B, info["dobj"] if variant == "full" else None, eps)
    if return_aux:
        return poses, {"Z": Z, "Q": Q, "info": info}
    return poses


def _stack(pts_2d, pts_3d, line_2d, line_3d, K):
    Cs, Ns = [], []
    if pts_2d is not None and len(pts_2d):
        C, N = point_constraints(pts_2d, pts_3d, K)
        Cs.append(C)
        Ns.append(N)
    if line_2d is not None and len(line_2d):
        C, N = line_constraints(line_2d, line_3d, K)
        Cs.append(C)
        Ns.append(N)
    return np.vstack(Cs), np.vstack(Ns)


def pnp(pts_2d, pts_3d, K, eps=1e-9, max_iters=2500, **kw):
    """cvxpnpl.py:523-552."""
    A, B = reduce_translation(*_stack(pts_2d, pts_3d, None, None, K))
    return solve_relaxation(A, B, eps, max_iters, **kw)


def pnl(line_2d, line_3d, K, eps=1e-9, max_iters=2500, **kw):
    """cvxpnpl.py:555-583."""
    A, B = reduce_translation(*_stack(None, None, line_2d, line_3d, K))
    return solve_relaxation(A, B, eps, max_iters, **kw)


def pnpl(pts_2d, line_2d, pts_3d, line_3d, K, eps=1e-9, max_iters=2500, **kw):
    """cvxpnpl.py:586-627."""
    A, B = reduce_translation(*_stack(pts_2d, pts_3d, line_2d, line_3d, K))
    return solve_relaxation(A, B, eps, max_iters, **kw)
