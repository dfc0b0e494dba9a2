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
// np.roots is backward stable (companion-matrix eigenvalues), also for the badly scaled
// quartics this branch produces (leading coefficient 1e-5 of the others: one root at 1e5,
// three near 1e-2).  Here: Ferrari's factorisation into two quadratics gives four complex
// starting points, the Aberth-Ehrlich simultaneous iteration on the ORIGINAL polynomial
// then converges to all four roots to working precision (Ferrari alone loses the small
// roots of such a quartic to cancellation in the depressing shift).
struct Cplx {
    double re, im;
};
CVX_HD Cplx c_add(Cplx a, Cplx b) { return Cplx{a.re + b.re, a.im + b.im}; }
CVX_HD Cplx c_sub(Cplx a, Cplx b) { return Cplx{a.re - b.re, a.im - b.im}; }
CVX_HD Cplx c_mul(Cplx a, Cplx b) { return Cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
CVX_HD Cplx c_div(Cplx a, Cplx b)
{
    // Smith's algorithm (no overflow of |b|^2)
    if (fabs(b.re) >= fabs(b.im)) {
        const double r = b.im / b.re, den = b.re + b.im * r;
        return Cplx{(a.re + a.im * r) / den, (a.im - a.re * r) / den};
    }
    const double r = b.re / b.im, den = b.re * r + b.im;
    return Cplx{(a.re * r + a.im) / den, (a.im * r - a.re) / den};
}
CVX_HD double c_abs1(Cplx a) { return fabs(a.re) + fabs(a.im); }

// Aberth-Ehrlich on a degree-n polynomial (ascending coefficients c[0..n], c[n] != 0), n <= 4
CVX_HD void aberth_polish(const double* c, int n, Cplx* z)
{
    for (int it = 0; it < 80; ++it) {
        Cplx w[4];
        double worst = 0.0;
        for (int k = 0; k < n; ++k) {
            // p(z_k), p'(z_k) by Horner; for |z| > 1 on the reversed polynomial (no overflow, and the
            // Newton ratio of a huge root keeps its accuracy): p(z) = z^n q(1/z)
            Cplx nr;   // Newton ratio p / p'
            if (c_abs1(z[k]) <= 1.0) {
                Cplx pv{c[n], 0.0}, dv{0.0, 0.0};
                for (int e = n - 1; e >= 0; --e) {
                    dv = c_add(c_mul(dv, z[k]), pv);
                    pv = c_add(c_mul(pv, z[k]), Cplx{c[e], 0.0});
                }
                if (pv.re == 0.0 && pv.im == 0.0) { w[k] = Cplx{0.0, 0.0}; continue; }
                nr = c_div(pv, dv);
            } else {
                const Cplx u = c_div(Cplx{1.0, 0.0}, z[k]);
                Cplx qv{c[0], 0.0}, dq{0.0, 0.0};
                for (int e = 1; e <= n; ++e) {
                    dq = c_add(c_mul(dq, u), qv);
                    qv = c_add(c_mul(qv, u), Cplx{c[e], 0.0});
                }
                if (qv.re == 0.0 && qv.im == 0.0) { w[k] = Cplx{0.0, 0.0}; continue; }
                // p'/p = n/z - u^2 q'(u)/q(u)  =>  p/p' = 1 / (n u - u^2 q'/q)
                const Cplx t = c_sub(Cplx{(double)n * u.re, (double)n * u.im}, c_mul(c_mul(u, u), c_div(dq, qv)));
                nr = c_div(Cplx{1.0, 0.0}, t);
            }
            Cplx sum{0.0, 0.0};
            for (int j = 0; j < n; ++j) {
                if (j == k) continue;
                Cplx df = c_sub(z[k], z[j]);
                if (df.re == 0.0 && df.im == 0.0) df = Cplx{1e-18 * (1.0 + c_abs1(z[k])), 1e-18 * (1.0 + c_abs1(z[k]))};
                sum = c_add(sum, c_div(Cplx{1.0, 0.0}, df));
            }
            const Cplx den = c_sub(Cplx{1.0, 0.0}, c_mul(nr, sum));
            w[k] = (den.re == 0.0 && den.im == 0.0) ? nr : c_div(nr, den);
            if (!isfinite(w[k].re) || !isfinite(w[k].im)) w[k] = Cplx{0.0, 0.0};
            worst = fmax(worst, c_abs1(w[k]) / fmax(c_abs1(z[k]), 1e-300));
        }
        for (int k = 0; k < n; ++k) z[k] = c_sub(z[k], w[k]);
        if (worst <= 4e-16) break;
    }
}

// the two roots of x^2 + b x + c as complex numbers
CVX_HD void quadratic_roots(double b, double c, Cplx& r0, Cplx& r1)
{
    const double disc = b * b - 4 * c;
    if (disc >= 0) {
        const double t = -0.5 * (b + copysign(sqrt(disc), b));
        r0 = Cplx{t, 0.0};
        r1 = Cplx{(t != 0.0) ? c / t : 0.0, 0.0};
    } else {
        r0 = Cplx{-0.5 * b, 0.5 * sqrt(-disc)};
        r1 = Cplx{-0.5 * b, -0.5 * sqrt(-disc)};
    }
}

// Returns the number of roots (4, or fewer when leading coefficients vanish exactly, like np.roots).
CVX_HD int quartic_real_parts(const double c[5], double x[4])
{
    int n = 4;
    while (n > 0 && c[n] == 0.0) --n;
    if (n == 0) return 0;
    Cplx z[4];
    if (n == 1) {
        x[0] = -c[0] / c[1];
        return 1;
    }
    if (n == 2) {
        quadratic_roots(c[1] / c[2], c[0] / c[2], z[0], z[1]);
    } else if (n == 3) {
        // cubic: one real root by Newton from outside the root bound, then the quadratic factor
        const double a = c[2] / c[3], b = c[1] / c[3], d = c[0] / c[3];
        double r = 1.0 + fmax(fabs(a), fmax(fabs(b), fabs(d)));
        for (int it = 0; it < 200; ++it) {
            const double f = ((r + a) * r + b) * r + d, fp = (3 * r + 2 * a) * r + b;
            const double dr = f / fp;
            r -= dr;
            if (!(fabs(dr) > 1e-16 * fabs(r))) break;
        }
        z[0] = Cplx{r, 0.0};
        const double qb = a + r;
        quadratic_roots(qb, b + r * qb, z[1], z[2]);
    } else {
        const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
        const double sh = 0.25 * a;
        // depressed quartic y^4 + p y^2 + q y + r, x = y - a/4
        const double p = b - 6 * sh * sh;
        const double q = cc - 2 * b * sh + 8 * sh * sh * sh;
        const double r = d - cc * sh + b * sh * sh - 3 * sh * sh * sh * sh;
        // resolvent cubic m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, largest real root (>= 0)
        const double A = p, Bc = 0.25 * p * p - r, Cc = -0.125 * q * q;
        double m = 1.0 + fmax(fabs(A), fmax(fabs(Bc), fabs(Cc)));
        for (int it = 0; it < 300; ++it) {
            const double f = ((m + A) * m + Bc) * m + Cc, fp = (3 * m + 2 * A) * m + Bc;
            const double dm = f / fp;
            m -= dm;
            if (!(fabs(dm) > 1e-16 * fabs(m))) break;
        }
        if (!(m > 0.0) || !isfinite(m)) m = 0.0;
        if (m > 0.0) {
            const double s2m = sqrt(2 * m), h = q / (2 * s2m);
            // y^2 - s2m y + (p/2 + m + h) = 0   and   y^2 + s2m y + (p/2 + m - h) = 0
            quadratic_roots(-s2m, 0.5 * p + m + h, z[0], z[1]);
            quadratic_roots(s2m, 0.5 * p + m - h, z[2], z[3]);
        } else {
            // biquadratic: y^2 = u with u^2 + p u + r = 0
            Cplx u0, u1;
            quadratic_roots(p, r, u0, u1);
            const Cplx us[2] = {u0, u1};
            for (int k = 0; k < 2; ++k) {
                // principal complex square root
                const double mod = sqrt(us[k].re * us[k].re + us[k].im * us[k].im);
                const double sr = sqrt(fmax(0.5 * (mod + us[k].re), 0.0));
                const double si = copysign(sqrt(fmax(0.5 * (mod - us[k].re), 0.0)), us[k].im);
                z[2 * k] = Cplx{sr, si};
                z[2 * k + 1] = Cplx{-sr, -si};
            }
        }
        for (int k = 0; k < 4; ++k) z[k].re -= sh;
    }
    // separate coincident starting points (a double root would make the Aberth weights singular)
    for (int k = 0; k < n; ++k) {
        if (!isfinite(z[k].re) || !isfinite(z[k].im)) z[k] = Cplx{0.3 * (k + 1), 0.2 * (k + 1)};
        const double pert = 1e-9 * (k + 1) * (1.0 + c_abs1(z[k]));
        z[k].re += (k & 1) ? pert : -pert;
        z[k].im += (k & 2) ? pert : -pert;
    }
    aberth_polish(c, n, z);
    for (int k = 0; k < n; ++k) x[k] = z[k].re;
    return n;
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
        bool fin = true;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            rc[i] = V[i * 10 + jmax] * inv;
            fin = fin && isfinite(rc[i]);
        }
        if (!fin) {   // np.linalg.svd raises LinAlgError on a non-finite matrix (cvxpnpl.py:510)
            status = ST_SINGULAR;
            return 0;
        }
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
        // a candidate that is not finite (e.g. the averaged quadratic of the rank-2 branch with a
        // vanishing leading coefficient, cvxpnpl.py:307-314) makes the reference's batched
        // np.linalg.svd raise LinAlgError for the whole call (cvxpnpl.py:510)
        for (int c = 0; c < 9 * n; ++c)
            if (!isfinite(rcs[c])) {
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
