// pnpl_extract.cuh -- pose recovery from the eigen-decomposition of the SDP
// solution (cvxpnpl.py:499-520), including the multi-solution branches
// _constraint_ortho_det (221-343) and _re6q3 (156-218).
//
// One thread per problem.  The rank-1 path works from registers; the rare rank>1
// paths use thread-local arrays (they may live in local memory -- they are off
// the hot loop).
#pragma once

#include "pnpl_core.cuh"

namespace cvx {

// ---- real parts of the four roots of  c4 x^4 + c3 x^3 + c2 x^2 + c1 x + c0 ------
// (cvxpnpl.py:185-186: np.roots, then np.real of ALL roots, complex ones included)
// Ferrari factorisation into two real quadratics + Newton polish of real roots.
// Returns the number of roots (4, or fewer when leading coefficients vanish).
CVX_HD int quartic_real_parts(const double c[5], double x[4])
{
    // strip vanishing leading coefficients like np.roots does for exact zeros
    if (c[4] == 0.0) {
        if (c[3] == 0.0) {
            if (c[2] == 0.0) {
                if (c[1] == 0.0) return 0;
                x[0] = -c[0] / c[1];
                return 1;
            }
            const double disc = c[1] * c[1] - 4 * c[2] * c[0];
            if (disc >= 0) {
                const double s = sqrt(disc);
                x[0] = (-c[1] + s) / (2 * c[2]);
                x[1] = (-c[1] - s) / (2 * c[2]);
            } else {
                x[0] = x[1] = -c[1] / (2 * c[2]);
            }
            return 2;
        }
        // cubic: deflate one real root found by Newton from outside the root bound
        const double a = c[2] / c[3], b = c[1] / c[3], d = c[0] / c[3];
        double r = 1.0 + fmax(fabs(a), fmax(fabs(b), fabs(d)));
        for (int it = 0; it < 200; ++it) {
            const double f = ((r + a) * r + b) * r + d, fp = (3 * r + 2 * a) * r + b;
            const double dr = f / fp;
            r -= dr;
            if (fabs(dr) <= 1e-16 * fabs(r)) break;
        }
        x[0] = r;
        const double qb = a + r, qc = b + r * qb;  // x^2 + qb x + qc
        const double disc = qb * qb - 4 * qc;
        if (disc >= 0) {
            const double s = sqrt(disc);
            x[1] = 0.5 * (-qb + s);
            x[2] = 0.5 * (-qb - s);
        } else {
            x[1] = x[2] = -0.5 * qb;
        }
        return 3;
    }
    const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
    const double sh = 0.25 * a;
    // depressed quartic y^4 + p y^2 + q y + r, x = y - a/4
    const double p = b - 6 * sh * sh;
    const double q = cc - 2 * b * sh + 8 * sh * sh * sh;
    const double r = d - cc * sh + b * sh * sh - 3 * sh * sh * sh * sh;
    double y[4];
    bool is_real[4];
    const double scale = fmax(fmax(fabs(p), sqrt(fabs(r))), 1e-300);
    if (fabs(q) <= 1e-14 * scale * sqrt(scale)) {
        // biquadratic: y^2 = (-p +- sqrt(p^2 - 4r)) / 2
        const double disc = p * p - 4 * r;
        if (disc >= 0) {
            const double s = sqrt(disc);
            const double u[2] = {0.5 * (-p + s), 0.5 * (-p - s)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (u[k] >= 0) {
                    y[2 * k] = sqrt(u[k]); y[2 * k + 1] = -sqrt(u[k]);
                    is_real[2 * k] = is_real[2 * k + 1] = true;
                } else {
                    y[2 * k] = y[2 * k + 1] = 0.0;  // purely imaginary pair
                    is_real[2 * k] = is_real[2 * k + 1] = false;
                }
            }
        } else {
            // y^2 complex: y = +-sqrt(u), u = (-p +- i sqrt(-disc))/2 ; Re sqrt(u) = sqrt((|u| + Re u)/2)
            const double re = -0.5 * p, im = 0.5 * sqrt(-disc);
            const double mod = sqrt(re * re + im * im);
            const double sr = sqrt(fmax(0.5 * (mod + re), 0.0));
            y[0] = sr; y[1] = sr; y[2] = -sr; y[3] = -sr;
            is_real[0] = is_real[1] = is_real[2] = is_real[3] = false;
        }
    } else {
        // resolvent cubic 8 m^3 + 8 p m^2 + (2 p^2 - 8 r) m - q^2 = 0, largest (positive) root
        const double A = p, Bc = 0.25 * p * p - r, Cc = -0.125 * q * q;  // m^3 + A m^2 + Bc m + Cc
        double m = 1.0 + fmax(fabs(A), fmax(fabs(Bc), fabs(Cc)));
        for (int it = 0; it < 300; ++it) {
            const double f = ((m + A) * m + Bc) * m + Cc, fp = (3 * m + 2 * A) * m + Bc;
            const double dm = f / fp;
            m -= dm;
            if (fabs(dm) <= 1e-16 * fabs(m)) break;
        }
        m = fmax(m, 1e-300);
        const double s2m = sqrt(2 * m);
        const double h = q / (2 * s2m);
        // y^2 - s2m y + (p/2 + m + h) = 0   and   y^2 + s2m y + (p/2 + m - h) = 0
        const double lin[2] = {-s2m, s2m};
        const double con[2] = {0.5 * p + m + h, 0.5 * p + m - h};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double disc = lin[k] * lin[k] - 4 * con[k];
            if (disc >= 0) {
                const double s = sqrt(disc);
                // stable quadratic roots
                const double t = -0.5 * (lin[k] + copysign(s, lin[k]));
                y[2 * k] = t;
                y[2 * k + 1] = (t != 0.0) ? con[k] / t : 0.0;
                is_real[2 * k] = is_real[2 * k + 1] = true;
            } else {
                y[2 * k] = y[2 * k + 1] = -0.5 * lin[k];
                is_real[2 * k] = is_real[2 * k + 1] = false;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double xv = y[k] - sh;
        if (is_real[k]) {
            // Newton polish on the original polynomial (guarded)
            for (int it = 0; it < 3; ++it) {
                const double f = (((c[4] * xv + c[3]) * xv + c[2]) * xv + c[1]) * xv + c[0];
                const double fp = ((4 * c[4] * xv + 3 * c[3]) * xv + 2 * c[2]) * xv + c[1];
                if (fp == 0.0) break;
                const double dx = f / fp;
                if (!(fabs(dx) < 1e-3 * (fabs(xv) + 1e-3))) break;  // near-multiple root: keep Ferrari value
                xv -= dx;
            }
        }
        x[k] = xv;
    }
    return 4;
}

// solve the n x n system (row-major, in place) with m right-hand sides by Gaussian
// elimination with partial pivoting.  Returns false when a pivot vanishes.
template <int NN, int MM>
CVX_HD bool gauss_solve(double* A, double* Bm)
{
    for (int k = 0; k < NN; ++k) {
        int piv = k;
        double best = fabs(A[k * NN + k]);
        for (int i = k + 1; i < NN; ++i)
            if (fabs(A[i * NN + k]) > best) { best = fabs(A[i * NN + k]); piv = i; }
        if (!(best > 0.0) || !isfinite(best)) return false;
        if (piv != k) {
            for (int j = 0; j < NN; ++j) { double t = A[k * NN + j]; A[k * NN + j] = A[piv * NN + j]; A[piv * NN + j] = t; }
            for (int j = 0; j < MM; ++j) { double t = Bm[k * MM + j]; Bm[k * MM + j] = Bm[piv * MM + j]; Bm[piv * MM + j] = t; }
        }
        const double ip = 1.0 / A[k * NN + k];
        for (int i = k + 1; i < NN; ++i) {
            const double f = A[i * NN + k] * ip;
            if (f == 0.0) continue;
            for (int j = k; j < NN; ++j) A[i * NN + j] -= f * A[k * NN + j];
            for (int j = 0; j < MM; ++j) Bm[i * MM + j] -= f * Bm[k * MM + j];
        }
    }
    for (int k = NN - 1; k >= 0; --k) {
        const double ip = 1.0 / A[k * NN + k];
        for (int j = 0; j < MM; ++j) {
            double s = Bm[k * MM + j];
            for (int i = k + 1; i < NN; ++i) s -= A[k * NN + i] * Bm[i * MM + j];
            Bm[k * MM + j] = s * ip;
        }
    }
    return true;
}

// The 21 quadratic forms of cvxpnpl.py:238-301 for the basis Vb (9 x k, row-major
// with leading dimension 4): form f -> symmetric k x k matrix P (row-major 4x4).
//   f in [0,6):   column products  c_i.c_j - delta_ij   (i<=j)
//   f in [6,12):  row products     r_i.r_j - delta_ij
//   f in [12,21): (c_i x c_j)_l - (c_kk)_l for (i,j,kk) cyclic, l = 0..2
// r = Vb alpha with alpha[k-1] = 1;  Vb row index 3*col + row.
CVX_HD_NOINLINE void quad_form(int f, const double* Vb, int k, double* P)
{
    for (int a = 0; a < 16; ++a) P[a] = 0.0;
    if (f < 12) {
        const bool cols = f < 6;
        const int g = cols ? f : f - 6;
        // g -> (i,j) with i<=j in the order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
        const int i = g < 3 ? 0 : (g < 5 ? 1 : 2);
        const int j = g < 3 ? g : (g < 5 ? g - 2 : 2);
        for (int m = 0; m < 3; ++m) {
            const double* vi = cols ? Vb + 4 * (3 * i + m) : Vb + 4 * (3 * m + i);
            const double* vj = cols ? Vb + 4 * (3 * j + m) : Vb + 4 * (3 * m + j);
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) P[4 * a + b] += vi[a] * vj[b];
        }
        if (i == j) P[4 * (k - 1) + (k - 1)] -= 1.0;
    } else {
        const int g = f - 12;
        const int i = g / 3, l = g % 3;
        const int j = (i + 1) % 3, kk = (i + 2) % 3;
        const int m = (l + 1) % 3, n = (l + 2) % 3;  // eps[l][m][n] = +1, eps[l][n][m] = -1
        const double* cim = Vb + 4 * (3 * i + m);
        const double* cin = Vb + 4 * (3 * i + n);
        const double* cjm = Vb + 4 * (3 * j + m);
        const double* cjn = Vb + 4 * (3 * j + n);
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) P[4 * b + a] += cim[a] * cjn[b] - cin[a] * cjm[b];
        const double* ck = Vb + 4 * (3 * kk + l);
        for (int b = 0; b < k; ++b) P[4 * (k - 1) + b] -= ck[b];
    }
    // symmetrise
    for (int a = 0; a < k; ++a)
        for (int b = a + 1; b < k; ++b) {
            const double s = 0.5 * (P[4 * a + b] + P[4 * b + a]);
            P[4 * a + b] = P[4 * b + a] = s;
        }
}

// rank-4 branch: the E6Q3 solver (cvxpnpl.py:156-218) on the normal equations of
// the 21 x 10 monomial matrix.  alpha_out: up to 4 x 4.  Returns #candidates, or
// -1 on a singular system (LinAlgError in the reference).
CVX_HD_NOINLINE int e6q3_candidates(const double* Vb, double* alpha_out)
{
    double BtB[36], BtC[24];
    for (int i = 0; i < 36; ++i) BtB[i] = 0.0;
    for (int i = 0; i < 24; ++i) BtC[i] = 0.0;
    for (int f = 0; f < 21; ++f) {
        double P[16];
        quad_form(f, Vb, 4, P);
        // monomials [a^2, b^2, c^2, ab, ac, bc | a, b, c, 1]   (cvxpnpl.py:318-331)
        const double row[10] = {P[0], P[5], P[10], 2 * P[1], 2 * P[2], 2 * P[6], 2 * P[3], 2 * P[7], 2 * P[11], P[15]};
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) BtB[6 * i + j] += row[i] * row[j];
            for (int j = 0; j < 4; ++j) BtC[4 * i + j] += row[i] * row[6 + j];
        }
    }
    if (!gauss_solve<6, 4>(BtB, BtC)) return -1;
    // D = -X[[1,2,5]]  (b^2, c^2, bc expressed through (a, b, c, 1))
    const double d00 = -BtC[4], d01 = -BtC[5], d02 = -BtC[6], d03 = -BtC[7];
    const double d10 = -BtC[8], d11 = -BtC[9], d12 = -BtC[10], d13 = -BtC[11];
    const double d20 = -BtC[20], d21 = -BtC[21], d22 = -BtC[22], d23 = -BtC[23];
    // Hidden-variable matrix M(a) = M0 + a M1 + a^2 M2 over (b, c, 1), from the
    // identities b(bc) = c(b^2), c(bc) = b(c^2), (bc)^2 = b^2 c^2 reduced once more
    // through the three relations (equals minus the reference's M, lines 190-202).
    double M0[9], M1[9];
    M0[0] = -d02 * d11 + d21 * d22 + d23;
    M0[1] = -d01 * d22 - d02 * d12 + d02 * d21 - d03 + d22 * d22;
    M0[2] = -d01 * d23 - d02 * d13 + d03 * d21 + d22 * d23;
    M0[3] = -d01 * d11 + d11 * d22 - d12 * d21 - d13 + d21 * d21;
    M0[4] = M0[0];
    M0[5] = -d03 * d11 - d12 * d23 + d13 * d22 + d21 * d23;
    M0[6] = -d01 * d01 * d11 - d01 * d12 * d21 - d01 * d13 + d01 * d21 * d21 - d02 * d11 * d12
            - d02 * d11 * d21 - d03 * d11 + d11 * d22 * d22 + 2 * d21 * d21 * d22 + 2 * d21 * d23;
    M0[7] = -d01 * d02 * d11 - d01 * d12 * d22 - d02 * d11 * d22 - d02 * d12 * d12 - d02 * d13
            + d02 * d21 * d21 - d03 * d12 + d12 * d22 * d22 + 2 * d21 * d22 * d22 + 2 * d22 * d23;
    M0[8] = -d01 * d03 * d11 - d01 * d12 * d23 - d02 * d11 * d23 - d02 * d12 * d13 - d03 * d13
            + d03 * d21 * d21 + d13 * d22 * d22 + 2 * d21 * d22 * d23 + d23 * d23;
    M1[0] = d20;
    M1[1] = -d00;
    M1[2] = d00 * d21 - d01 * d20 - d02 * d10 + d20 * d22;
    M1[3] = -d10;
    M1[4] = d20;
    M1[5] = -d00 * d11 + d10 * d22 - d12 * d20 + d20 * d21;
    M1[6] = -d00 * d11 - d01 * d10 + 2 * d20 * d21;
    M1[7] = -d00 * d12 - d02 * d10 + 2 * d20 * d22;
    M1[8] = -d00 * d01 * d11 - d00 * d13 + d00 * d21 * d21 - d01 * d12 * d20 - d02 * d10 * d12
            - d02 * d11 * d20 - d03 * d10 + d10 * d22 * d22 + 2 * d20 * d21 * d22 + 2 * d20 * d23;
    const double M2_22 = -d00 * d10 + d20 * d20;

    // det M(a): rows 0,1 are affine in a, row 2 has degree <= 2 (only [2][2]).
    // 2x2 minors of rows 0,1 (degree 2), columns (c1,c2):
    double mn[3][3];  // mn[col removed][power]
#pragma unroll
    for (int rm = 0; rm < 3; ++rm) {
        const int c1 = rm == 0 ? 1 : 0, c2 = rm == 2 ? 1 : 2;
        const double a0 = M0[c1], a1 = M1[c1], b0 = M0[c2], b1 = M1[c2];
        const double e0 = M0[3 + c1], e1 = M1[3 + c1], f0 = M0[3 + c2], f1 = M1[3 + c2];
        // (a0 + a1 x)(f0 + f1 x) - (b0 + b1 x)(e0 + e1 x)
        mn[rm][0] = a0 * f0 - b0 * e0;
        mn[rm][1] = a0 * f1 + a1 * f0 - b0 * e1 - b1 * e0;
        mn[rm][2] = a1 * f1 - b1 * e1;
    }
    // det = M[2][0]*mn[0] - M[2][1]*mn[1] + M[2][2]*mn[2]
    double q[5] = {0, 0, 0, 0, 0};  // ascending powers
    const double r0[3] = {M0[6], M1[6], 0.0}, r1[3] = {M0[7], M1[7], 0.0}, r2[3] = {M0[8], M1[8], M2_22};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (i + j > 4) continue;
            q[i + j] += r0[i] * mn[0][j] - r1[i] * mn[1][j] + r2[i] * mn[2][j];
        }
    double roots[4];
    const int nr = quartic_real_parts(q, roots);
    for (int k = 0; k < nr; ++k) {
        const double a = roots[k];
        double L[6], rhs[3];  // L: 3x2
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            L[2 * r] = M0[3 * r] + a * M1[3 * r];
            L[2 * r + 1] = M0[3 * r + 1] + a * M1[3 * r + 1];
            rhs[r] = M0[3 * r + 2] + a * M1[3 * r + 2] + (r == 2 ? a * a * M2_22 : 0.0);
        }
        double G[4] = {L[0] * L[0] + L[2] * L[2] + L[4] * L[4], L[0] * L[1] + L[2] * L[3] + L[4] * L[5], 0,
                       L[1] * L[1] + L[3] * L[3] + L[5] * L[5]};
        G[2] = G[1];
        double g[2] = {L[0] * rhs[0] + L[2] * rhs[1] + L[4] * rhs[2], L[1] * rhs[0] + L[3] * rhs[1] + L[5] * rhs[2]};
        if (!gauss_solve<2, 1>(G, g)) return -1;
        alpha_out[4 * k] = a;
        alpha_out[4 * k + 1] = -g[0];
        alpha_out[4 * k + 2] = -g[1];
        alpha_out[4 * k + 3] = 1.0;
    }
    return nr;
}

// cvxpnpl.py:221-343.  order[]: eigenvector indices by ascending eigenvalue.
// rc_out: up to 4 x 9.  Returns the number of candidates, -1 singular.
template <int S>
CVX_HD_NOINLINE int multi_solution(Arr<S> V, const int order[10], int rank, double* rc_out)
{
    const int k = (rank <= 2) ? 2 : 4;  // min(ceil(rank/2)*2, 4)
    // rows = top-k eigenvectors in ascending order; de-homogenise (lines 234-236)
    double Vb[36];  // 9 x k, leading dimension 4
    {
        const int jl = order[9];
        const double inv = 1.0 / V[90 + jl];
        for (int a = 0; a < k - 1; ++a) {
            const int ja = order[10 - k + a];
            const double w = V[90 + ja] * inv;
            for (int i = 0; i < 9; ++i) Vb[4 * i + a] = V[i * 10 + ja] - w * V[i * 10 + jl];
        }
        for (int i = 0; i < 9; ++i) Vb[4 * i + (k - 1)] = V[i * 10 + jl] * inv;
        for (int a = k; a < 4; ++a)
            for (int i = 0; i < 9; ++i) Vb[4 * i + a] = 0.0;
    }
    double alpha[16];
    int n;
    if (k == 2) {
        // lines 303-315: average the 21 quadratics in a
        double c0 = 0, c1 = 0, c2 = 0;
        for (int f = 0; f < 21; ++f) {
            double P[16];
            quad_form(f, Vb, 2, P);
            c0 += P[0];
            c1 += 2 * P[1];
            c2 += P[5];
        }
        c0 *= (1.0 / 21.0); c1 *= (1.0 / 21.0); c2 *= (1.0 / 21.0);
        const double root = sqrt(fmax(c1 * c1 - 4 * c0 * c2, 0.0));
        alpha[0] = (-c1 + root) / (2 * c0); alpha[1] = 1.0;
        alpha[4] = (-c1 - root) / (2 * c0); alpha[5] = 1.0;
        n = 2;
    } else {
        n = e6q3_candidates(Vb, alpha);
        if (n < 0) return -1;
    }
    for (int c = 0; c < n; ++c)
        for (int i = 0; i < 9; ++i) {
            double s = 0;
            for (int a = 0; a < k; ++a) s = fma(alpha[4 * c + a], Vb[4 * i + a], s);
            rc_out[9 * c + i] = s;
        }
    return n;
}

// cvxpnpl.py:493-520.  V/lam: eigen-decomposition of the final iterate (Z has
// eigenvalues max(lam, 0)).  Writes up to 4 poses (NaN padded); returns n_poses and
// updates status (NaN / singular / rank-0 / not-certified flag).
template <int S, class QIn, class BIn>
CVX_HD int extract_poses(Arr<S> V, const double lam[10], QIn Q, BIn Bm, int32_t& status, double dobj,
                         double eps, double* R_out, double* t_out, double& pobj)
{
    const double qnan = nan("");
#pragma unroll
    for (int i = 0; i < 36; ++i) R_out[i] = qnan;
#pragma unroll
    for (int i = 0; i < 12; ++i) t_out[i] = qnan;
    pobj = qnan;
    if ((status & 0xff) == ST_NAN) return 1;  // single NaN pose (cvxpnpl.py:498)

    int rank = 0, jmax = 0;
    double lmax = lam[0];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        rank += (lam[j] > 1e-3) ? 1 : 0;
        if (lam[j] > lmax) { lmax = lam[j]; jmax = j; }
    }
    int n;
    bool certified = true;
    if (rank == 1) {
        double rc[9];
        const double inv = 1.0 / V[90 + jmax];
#pragma unroll
        for (int i = 0; i < 9; ++i) rc[i] = V[i * 10 + jmax] * inv;
        pobj = finish_pose(rc, Q, Bm, R_out, t_out);
        if (eps >= 0) certified = !(fabs(pobj - dobj) > eps);
        n = 1;
    } else if (rank == 0) {
        status = ST_RANK0;
        return 0;
    } else {
        // ascending order of eigenvalues (stable insertion sort on indices)
        double l[10];
        int order[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) l[j] = lam[j];
        for (int j = 0; j < 10; ++j) order[j] = j;
        for (int a = 1; a < 10; ++a) {
            const int ia = order[a];
            int b = a - 1;
            while (b >= 0 && l[order[b]] > l[ia]) { order[b + 1] = order[b]; --b; }
            order[b + 1] = ia;
        }
        double rcs[36];
        n = multi_solution(V, order, rank, rcs);
        if (n < 0) {
            status = ST_SINGULAR;
            return 0;
        }
        for (int c = 0; c < n; ++c) {
            const double o = finish_pose(rcs + 9 * c, Q, Bm, R_out + 9 * c, t_out + 3 * c);
            if (c == 0) pobj = o;
            if (eps >= 0 && fabs(o - dobj) > eps) certified = false;
        }
    }
    if (!certified) status |= ST_FLAG_NOT_CERTIFIED;
    return n;
}

}  // namespace cvx
