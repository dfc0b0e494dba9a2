// pnpl_dr.inl -- the inner loop of the SDP solver (Jacobi sweep, basis change, DR
// step), written once in terms of CVX_REAL and included twice by pnpl_core.cuh:
// as double in namespace cvx (the solver proper) and as float in namespace cvx::f32
// (the FP32 first phase that brings a problem into the linear tail, see
// pnpl_solve.cuh).  No include guard on purpose.
// ---------------------------------------------------------------------------------
// Jacobi symmetric eigensolver, register resident, compact code.
//
// t[55] is the packed 10x10 matrix held in REGISTERS (every index below is a
// compile-time constant); V (eigenbasis, V[i*10+j] = component i of eigenvector
// j) stays in the problem's strided shared-memory view.
//
// Pivot order: round-robin tournament, 9 rounds of 5 disjoint pairs.  To keep the
// loop body SMALL (the instruction cache, not the FP64 pipe, bounded the fully
// unrolled version: ncu stall_no_instruction 2.4 cycles per issue) every round
// rotates the same fixed position pairs (0,9) (1,8) (2,7) (3,6) (4,5) and then
// applies the fixed tournament permutation  0->0, i->i+1 (1..8), 9->1  to the
// rows/columns of t and to the columns of V ("the players move, the tables
// stay"), so one round body is executed nine times by a rolled loop.  The order
// of the eigenpairs is irrelevant to the caller (lam[j] always matches column j
// of V), and after 9 rounds the permutation is the identity again.
//
// The five rotations of a round commute and their angles only depend on entries
// no other rotation of the round touches, so all five (c, s) are computed up
// front (instruction-level parallelism across the sqrt / divide chains).
// ---------------------------------------------------------------------------------
// reciprocal square root: device intrinsic path / host libm
CVX_HD CVX_REAL cvx_rsqrt(CVX_REAL x)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return CVX_REAL(1.0) / (CVX_REAL)sqrt((double)x);
#endif
}

// Rotation for pivot (p,q) annihilating a_pq (classical Jacobi, |angle| <= pi/4):
//   d = a_qq - a_pp, b = 2 a_pq, h = sqrt(d^2 + b^2)
//   cos^2 = (h + |d|) / (2h),  sin = sgn(d) b / (2 h cos),  tan = sin / cos
// written with two reciprocal square roots and no division, which keeps the
// dependent chain short (the five chains of a round are the critical path).
CVX_HD void jacobi_cs(CVX_REAL app, CVX_REAL aqq, CVX_REAL apq, CVX_REAL& c, CVX_REAL& s, CVX_REAL& tn)
{
    const CVX_REAL d = aqq - app, b2 = CVX_REAL(2.0) * apq;
    const CVX_REAL g = fma(d, d, b2 * b2);
    const CVX_REAL ad = fabs(d);
    // negligible pivot (also covers d = b2 = 0 and underflow of g): identity rotation
    const bool skip = !(fabs(apq) > CVX_RELSKIP * ad) || !(g > CVX_TINY);
    const CVX_REAL ig = cvx_rsqrt(skip ? CVX_REAL(1.0) : g);
    const CVX_REAL c2 = fma(CVX_REAL(0.5) * ad, ig, CVX_REAL(0.5));          // in [0.5, 1]
    const CVX_REAL rc = cvx_rsqrt(c2);
    const CVX_REAL sg = copysign(CVX_REAL(0.5), d) * b2 * ig;       // sin * cos
    c = skip ? CVX_REAL(1.0) : c2 * rc;
    s = skip ? CVX_REAL(0.0) : sg * rc;
    tn = skip ? CVX_REAL(0.0) : sg * rc * rc;
}

// one sweep; returns the off-diagonal square sum seen at the pivots (before they
// are annihilated)
template <int S>
CVX_HD CVX_REAL jacobi_sweep_reg(CVX_REAL t[55], ArrT<S, CVX_REAL> V)
{
#if defined(CVX_STALE_ANGLES) && !defined(__CUDA_ARCH__)
    // EXPERIMENT (host build only): all 45 rotation angles from the matrix as it is at the START of the sweep, then
    // the nine rounds applied without recomputing them -- the nine dependent angle chains of a sweep become one.
    {
        CVX_REAL t0[55], off0 = CVX_REAL(0.0);
        for (int e = 0; e < 55; ++e) t0[e] = t[e];
        int perm[10];
        for (int i = 0; i < 10; ++i) perm[i] = i;
        for (int round = 0; round < 9; ++round) {
            CVX_REAL cs[5], sn[5], tn[5];
            for (int k = 0; k < 5; ++k) {
                const int p = jp_p(k), q = jp_q(k);
                const int op = perm[p], oq = perm[q];
                const CVX_REAL apq0 = t0[op > oq ? sidx(op, oq) : sidx(oq, op)];
                off0 = fma(apq0, apq0, off0);
                jacobi_cs(t0[sidx(op, op)], t0[sidx(oq, oq)], apq0, cs[k], sn[k], tn[k]);
            }
            for (int k = 0; k < 5; ++k) {
                const int p = jp_p(k), q = jp_q(k);
                const CVX_REAL c = cs[k], s = sn[k];
                const CVX_REAL app = t[sidx(p, p)], aqq = t[sidx(q, q)], apq = t[sidx(q, p)];
                t[sidx(p, p)] = c * c * app - 2 * c * s * apq + s * s * aqq;
                t[sidx(q, q)] = s * s * app + 2 * c * s * apq + c * c * aqq;
                t[sidx(q, p)] = c * s * (app - aqq) + (c * c - s * s) * apq;
                for (int m = 0; m < 10; ++m) {
                    if (m == p || m == q) continue;
                    const CVX_REAL amp = t[sidx(m, p)], amq = t[sidx(m, q)];
                    t[sidx(m, p)] = fma(c, amp, -s * amq);
                    t[sidx(m, q)] = fma(s, amp, c * amq);
                }
            }
            {
                CVX_REAL u[55];
                int np[10];
                for (int i = 0; i < 10; ++i)
                    for (int j = 0; j <= i; ++j) u[sidx(jp_sigma(i), jp_sigma(j))] = t[sidx(i, j)];
                for (int e = 0; e < 55; ++e) t[e] = u[e];
                for (int i = 0; i < 10; ++i) np[jp_sigma(i)] = perm[i];
                for (int i = 0; i < 10; ++i) perm[i] = np[i];
            }
            for (int row = 0; row < 10; ++row) {
                CVX_REAL v[10];
                for (int j = 0; j < 10; ++j) v[j] = V[row * 10 + j];
                for (int k = 0; k < 5; ++k) {
                    const int p = jp_p(k), q = jp_q(k);
                    const CVX_REAL vp = v[p], vq = v[q];
                    v[p] = fma(cs[k], vp, -sn[k] * vq);
                    v[q] = fma(sn[k], vp, cs[k] * vq);
                }
                for (int j = 0; j < 10; ++j) V[row * 10 + jp_sigma(j)] = v[j];
            }
        }
        return off0;
    }
#endif
    CVX_REAL off = CVX_REAL(0.0);
#pragma unroll 1
    for (int round = 0; round < 9; ++round) {
        CVX_REAL cs[5], sn[5], tn[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int p = jp_p(k), q = jp_q(k);
            const CVX_REAL apq = t[sidx(q, p)];
            off = fma(apq, apq, off);
            jacobi_cs(t[sidx(p, p)], t[sidx(q, q)], apq, cs[k], sn[k], tn[k]);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int p = jp_p(k), q = jp_q(k);
            const CVX_REAL c = cs[k], s = sn[k];
            const CVX_REAL apq = t[sidx(q, p)];
            t[sidx(p, p)] = fma(-tn[k], apq, t[sidx(p, p)]);
            t[sidx(q, q)] = fma(tn[k], apq, t[sidx(q, q)]);
            t[sidx(q, p)] = CVX_REAL(0.0);
#pragma unroll
            for (int m = 0; m < 10; ++m) {
                if (m == p || m == q) continue;
                const CVX_REAL amp = t[sidx(m, p)], amq = t[sidx(m, q)];
                t[sidx(m, p)] = fma(c, amp, -s * amq);
                t[sidx(m, q)] = fma(s, amp, c * amq);
            }
        }
        // tournament permutation of rows/columns of t
        {
            CVX_REAL u[55];
#pragma unroll
            for (int i = 0; i < 10; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) u[sidx(jp_sigma(i), jp_sigma(j))] = t[sidx(i, j)];
#pragma unroll
            for (int e = 0; e < 55; ++e) t[e] = u[e];
        }
        // rotate + permute the columns of V, two rows at a time (two independent
        // instruction streams for the single resident warp of the scheduler)
#pragma unroll 1
        for (int row = 0; row < 10; row += 2) {
            CVX_REAL v[2][10];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 10; ++j) v[h][j] = V[(row + h) * 10 + j];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int p = jp_p(k), q = jp_q(k);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const CVX_REAL vp = v[h][p], vq = v[h][q];
                    v[h][p] = fma(cs[k], vp, -sn[k] * vq);
                    v[h][q] = fma(sn[k], vp, cs[k] * vq);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 10; ++j) V[(row + h) * 10 + jp_sigma(j)] = v[h][j];
        }
    }
    return off;
}

// T <- V' M V (packed, into the strided view T).  Rolled over blocks of two
// columns: w_j = M v_j with M read at compile-time offsets, then one dot product
// per (i, j) pair.  Compact code on purpose (see above).
template <int S>
CVX_HD void rotate_into_basis(ArrT<S, CVX_REAL> M, ArrT<S, CVX_REAL> V, ArrT<S, CVX_REAL> T)
{
#pragma unroll 1
    for (int j0 = 0; j0 < 10; j0 += 2) {
        CVX_REAL w0[10], w1[10];
        {
            CVX_REAL a0[10], a1[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                a0[k] = V[k * 10 + j0];
                a1[k] = V[k * 10 + j0 + 1];
                w0[k] = CVX_REAL(0.0);
                w1[k] = CVX_REAL(0.0);
            }
#pragma unroll
            for (int r = 0; r < 10; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    const CVX_REAL m = M[sidx(r, c)];
                    w0[r] = fma(m, a0[c], w0[r]);
                    w1[r] = fma(m, a1[c], w1[r]);
                    if (r != c) {
                        w0[c] = fma(m, a0[r], w0[c]);
                        w1[c] = fma(m, a1[r], w1[c]);
                    }
                }
        }
        // rows i = j0 .. 9 in pairs (j0 is even, so the count is even): four
        // independent dot-product chains
#pragma unroll 1
        for (int i = j0; i < 10; i += 2) {
            CVX_REAL s00 = CVX_REAL(0.0), s01 = CVX_REAL(0.0), s10 = CVX_REAL(0.0), s11 = CVX_REAL(0.0);
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                const CVX_REAL va = V[k * 10 + i], vb = V[k * 10 + i + 1];
                s00 = fma(va, w0[k], s00);
                s01 = fma(va, w1[k], s01);
                s10 = fma(vb, w0[k], s10);
                s11 = fma(vb, w1[k], s11);
            }
            const int ba = (i * (i + 1)) / 2 + j0, bb = ((i + 1) * (i + 2)) / 2 + j0;
            T[ba] = s00;
            // (i, j0+1) with i == j0 is the transposed duplicate of (j0+1, j0): skip
            if (i > j0) T[ba + 1] = s01;
            T[bb] = s10;
            T[bb + 1] = s11;
        }
    }
}

// memory-resident convenience wrapper (cold start on a matrix held in a strided
// view); used by the extraction stage kernel only.
template <int S>
CVX_HD CVX_REAL jacobi_sweep(ArrT<S, CVX_REAL> T, ArrT<S, CVX_REAL> V)
{
    CVX_REAL t[55];
#pragma unroll
    for (int e = 0; e < 55; ++e) t[e] = T[e];
    const CVX_REAL off = jacobi_sweep_reg(t, V);
#pragma unroll
    for (int e = 0; e < 55; ++e) T[e] = t[e];
    return off;
}

// ---------------------------------------------------------------------------------
// One Douglas-Rachford step of   min <Q,Z>  s.t.  Z in Affine (22 equalities of
// cvxpnpl.py:387-448) and Z in PSD:
//     Z  = P_psd(M)        (from the eigen-pairs lam, V of M)
//     X  = P_aff(2 Z - M - Q/rho)
//     M += alpha (X - Z)
// Returns ||X - Z||_F^2, the fixed-point residual (primal residual X-Z and dual
// residual rho (M+ - M)/alpha coincide up to scale).  Q/rho is read through `qr`
// (45 packed entries of the 9x9 block); the eigenvalues through the strided view
// L (10).  Also returns Z in z[] (registers) and writes the step g = alpha (X - Z)
// (already added to M) to the strided view G for the Anderson accelerator.
//
// P_aff in closed form: the 15 triples are mutually orthogonal, so each is fixed by
// subtracting its own normal component (the signed mean of its three entries when
// sigma = 1); the remaining 7 equalities (rank 6) only touch the diagonal:
// Z99 = 1 and the 3x3 array D[r][c] = Z[3c+r, 3c+r] has unit row and column sums.
//
// Homogeneous scaling (a diagonal preconditioner): the iteration runs on
// Z' = D Z D with D = diag(1,..,1,sigma), isig = 1/sigma.  The PSD cone is invariant
// under the congruence, Q' = Q (its last row/column is zero), Z'99 = sigma^2 and the
// nine triples that touch row 9 pick up the coefficient 1/sigma on that entry.
// sigma ~ 1.5 cuts the iteration count by a third on PnP/PnPL.
//
// rowk = 1 is the reference's SDP (22 equalities); rowk = 0 is the "rc" ablation of
// benchmarks/toolkit/methods/rc.py:9-60 (the six row-orthonormality equalities removed).
// ---------------------------------------------------------------------------------
// Second half of a DR step: given Z = P_psd(M) through the accessor zf(i, j) (i >= j), X = P_aff(2 Z - M - Q/rho),
// M += alpha (X - Z), G = alpha (X - Z); returns ||X - Z||_F^2.  The accessor is either an array in registers
// (ZPacked: the full-decomposition solver) or the rank-two form of the tracked solver (ZRank2: two multiply-adds
// per entry instead of 55 live doubles).
struct ZPacked {
    const CVX_REAL* z;
    CVX_HD CVX_REAL operator()(int i, int j) const { return z[sidx(i, j)]; }
};
struct ZRank2 {
    CVX_REAL a[10], ta[10], b[10], tb[10];   // Z = (t0 a) a' + (t1 b) b'
    CVX_HD CVX_REAL operator()(int i, int j) const { return fma(ta[i], a[j], tb[i] * b[j]); }
};
// PART 2: everything; PART 0: the first eight triples; PART 1: the other seven and the diagonal equalities (the two
// threads of a problem in the role-split solver, pnpl_track2.cuh; the return values add up).
template <int PART, int S, class QR, class ZF>
CVX_HD CVX_REAL dr_affine_part(ArrT<S, CVX_REAL> M, ArrT<S, CVX_REAL> G, QR qr, CVX_REAL alpha, CVX_REAL isig, CVX_REAL rowk,
                               const ZF& zf)
{
    CVX_REAL res = CVX_REAL(0.0);
    const CVX_REAL inrm9 = CVX_REAL(1.0) / (CVX_REAL(2.0) + isig * isig);
#define CVX_Q(i, j) (((i) < 9 && (j) < 9) ? qr[sidx(i, j)] : CVX_REAL(0.0))
#define CVX_TRI(i0, j0, s0, i1, j1, s1, i2, j2, s2, ROW)                                    \
    {                                                                                        \
        const int e0 = sidx(i0, j0), e1 = sidx(i1, j1), e2 = sidx(i2, j2);                  \
        const CVX_REAL m0 = M[e0], m1 = M[e1], m2 = M[e2];                                    \
        const CVX_REAL z0 = zf(i0, j0), z1 = zf(i1, j1), z2 = zf(i2, j2);                     \
        const CVX_REAL w0 = CVX_REAL(2.0) * z0 - m0 - CVX_Q(i0, j0);                                 \
        const CVX_REAL w1 = CVX_REAL(2.0) * z1 - m1 - CVX_Q(i1, j1);                                 \
        const CVX_REAL w2 = CVX_REAL(2.0) * z2 - m2 - CVX_Q(i2, j2);                                 \
        /* third entry on the homogeneous row carries 1/sigma in the scaled problem */      \
        const CVX_REAL a2 = ((i2) == 9) ? (s2) * isig : (CVX_REAL)(s2);                         \
        const CVX_REAL r = ((s0) * w0 + (s1) * w1 + a2 * w2) *                                 \
                         (((i2) == 9) ? inrm9 : ((ROW) ? rowk * (CVX_REAL(1.0) / CVX_REAL(3.0)) : (CVX_REAL(1.0) / CVX_REAL(3.0))));  \
        const CVX_REAL d0 = w0 - (s0) * r - z0;                                               \
        const CVX_REAL d1 = w1 - (s1) * r - z1;                                               \
        const CVX_REAL d2 = w2 - a2 * r - z2;                                                 \
        M[e0] = fma(alpha, d0, m0);                                                         \
        M[e1] = fma(alpha, d1, m1);                                                         \
        M[e2] = fma(alpha, d2, m2);                                                         \
        G[e0] = alpha * d0;                                                                 \
        G[e1] = alpha * d1;                                                                 \
        G[e2] = alpha * d2;                                                                 \
        res += CVX_REAL(2.0) * (d0 * d0 + d1 * d1 + d2 * d2);                                         \
    }
    if (PART != 1) {
        CVX_TRIPLES_A(CVX_TRI)
    }
    if (PART != 0) {
        CVX_TRIPLES_B(CVX_TRI)
    }
#undef CVX_TRI
    // diagonal block
    if (PART != 0) {
        CVX_REAL w[9], md[10], zd[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) zd[i] = zf(i, i);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            md[i] = M[sidx(i, i)];
            w[i] = CVX_REAL(2.0) * zd[i] - md[i] - qr[sidx(i, i)];
        }
        md[9] = M[sidx(9, 9)];
        // D[r][c] = w[3c + r]; project onto unit row sums (over c) and column sums (over r)
        CVX_REAL R[3], C[3], Gs = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) R[r] = w[r] + w[3 + r] + w[6 + r];
#pragma unroll
        for (int c = 0; c < 3; ++c) { C[c] = w[3 * c] + w[3 * c + 1] + w[3 * c + 2]; Gs += C[c]; }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int i = 3 * c + r;
                // rowk = 0 ("rc" variant): only the column sums are constrained
                const CVX_REAL x = w[i] - rowk * (R[r] - CVX_REAL(1.0)) * (CVX_REAL(1.0) / CVX_REAL(3.0)) - (C[c] - CVX_REAL(1.0)) * (CVX_REAL(1.0) / CVX_REAL(3.0))
                                 + rowk * (Gs - CVX_REAL(3.0)) * (CVX_REAL(1.0) / CVX_REAL(9.0));
                const CVX_REAL d = x - zd[i];
                M[sidx(i, i)] = fma(alpha, d, md[i]);
                G[sidx(i, i)] = alpha * d;
                res = fma(d, d, res);
            }
        const CVX_REAL d9 = CVX_REAL(1.0) / (isig * isig) - zd[9];
        M[sidx(9, 9)] = fma(alpha, d9, md[9]);
        G[sidx(9, 9)] = alpha * d9;
        res = fma(d9, d9, res);
    }
#undef CVX_Q
    return res;
}

template <int S, class QR, class ZF>
CVX_HD CVX_REAL dr_affine_update(ArrT<S, CVX_REAL> M, ArrT<S, CVX_REAL> G, QR qr, CVX_REAL alpha, CVX_REAL isig, CVX_REAL rowk,
                                 const ZF& zf)
{
    return dr_affine_part<2>(M, G, qr, alpha, isig, rowk, zf);
}

template <int S, class QR>
CVX_HD CVX_REAL dr_step(ArrT<S, CVX_REAL> M, ArrT<S, CVX_REAL> V, ArrT<S, CVX_REAL> L, ArrT<S, CVX_REAL> G, QR qr, CVX_REAL alpha, CVX_REAL isig, CVX_REAL rowk,
                      CVX_REAL z[55])
{
#pragma unroll
    for (int e = 0; e < 55; ++e) z[e] = CVX_REAL(0.0);
#pragma unroll 1
    for (int j = 0; j < 10; ++j) {
        const CVX_REAL lj = L[j];
        if (lj > CVX_REAL(0.0)) {
            CVX_REAL v[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) v[k] = V[k * 10 + j];
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                const CVX_REAL lr = lj * v[r];
#pragma unroll
                for (int c = 0; c <= r; ++c) z[sidx(r, c)] = fma(lr, v[c], z[sidx(r, c)]);
            }
        }
    }
    return dr_affine_update(M, G, qr, alpha, isig, rowk, ZPacked{z});
}

