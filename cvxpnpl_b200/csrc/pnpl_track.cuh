// pnpl_track.cuh -- PSD projection by TRACKING the positive eigenpairs of the DR iterate.
//
// Why.  ncu on the round-1 solver: 55 % of the issue slots of a DR pass went into a full
// 10x10 eigen-decomposition (basis change V'MV + one cyclic Jacobi sweep, ~7.7 k of ~14 k
// instructions) although Z = P_psd(M) only needs the POSITIVE eigenpairs of M, and M has
// at most two of them from the fifth iteration on (measured on seeded batches, host build:
// PnPL 8+4 0 of 21 k iterations with three, PnP-8 0.07 %, PnL-6 1.4 %; the first four
// iterations have up to six).  So:
//   * the first iterations (Opts::early, 2 or 4) run with the full decomposition in the lane-parallel
//     pre-pass kernel (all lanes in the same phase: no divergence);
//   * from then on a problem carries two orthonormal vectors u0, u1 with Ritz values
//     th0, th1 -- the positive eigenpairs, or, for a slot whose value is not positive, a
//     sentinel that follows the eigenvalue closest to zero -- and every DR iteration
//     refines them for the new M by ONE step of block Rayleigh-quotient iteration in a
//     basis where the solve is an 8x8 Cholesky:
//       1. two Householder reflectors H1 H2 with (H2 H1) [u0 u1] = [+-e8 +-e9];
//          M' = H2 H1 M H1 H2 (two symmetric rank-2 updates, registers)
//       2. M' = [[B C][C' A]]: diagonalise the 2x2 block A (Ritz values th~_i, C~ = C J)
//       3. x_i = (s_i I - B)^-1 c~_i with s_i = th~_i for a positive slot (RQI: cubic) and
//          s_i = 0 for a sentinel slot (inverse iteration towards the eigenvalue nearest 0).
//          The Cholesky factorisations double as the CERTIFICATE that nothing else in the
//          spectrum is positive: s I - B > 0 for s = 0, or for both positive slots plus a
//          third factorisation at s = 0.  A non-positive pivot = an untracked positive
//          eigenvalue (or a slot that lost its eigenvector): the problem is handed to the
//          warp-per-problem kernel, which decomposes M in full (pnpl_warp.cuh).
//       4. new basis [x_i; e~_i]: 2x2 Rayleigh-Ritz from dot products of 8-vectors (B x_i =
//          s_i x_i - c~_i), back through the reflectors.
//     ~1.4 k FMA + ~30 rsqrt per iteration instead of ~5 k + 45 rotation angles, the same
//     exact projection (error cubic in the per-iteration change of M), and a per-problem
//     state of 20 instead of 110 doubles for the decomposition.
// Everything here is __host__ __device__ like the rest of the per-problem code; the host
// build (tests/host) runs the same routines.
#pragma once

#include "pnpl_core.cuh"
#include "pnpl_solve.cuh"

namespace cvx {

// DR iterations with the full decomposition (pre-pass) before a problem is tracked: Opts::early.  The first iterates
// have up to six positive eigenvalues.  Hand-backs (host build, 1000 problems each, tolerance 8..12) with 4 / 3 / 2 / 1
// early iterations: PnPL 8+4 0 / 0 / 0 / 5 %, PnP-8 0.2 / 1.4 / 1.5 / 14 %, PnL-6 4 / 7 / 10 / 50 %: two are enough for
// the well-constrained point + line problems, four otherwise.
#ifndef CVX_TRK_EARLY
#define CVX_TRK_EARLY 0   // > 0: override for experiments
#endif
CVX_HD int default_early(int n_pts, int n_lines)
{
    if (CVX_TRK_EARLY > 0) return CVX_TRK_EARLY;
    return (n_pts >= 8 && n_lines >= 4) ? 2 : 4;
}

#if defined(CVX_TRK_DEBUG) && !defined(__CUDA_ARCH__)
static long g_track_steps = 0, g_track_passes = 0;
#endif
// track_step / pass result bits
enum : int {
    TRK_OK = 0,
    TRK_NEED_FULL = 1,   // certificate failed: an untracked eigenvalue is positive, or a slot lost its vector
};

// Symmetric two-sided Householder update  m <- (I - beta v v') m (I - beta v v')  on the packed
// lower triangle held in registers; v has N leading non-zeros (v[N..9] = 0).
template <int N>
CVX_HD void reflect_sym(double m[55], const double v[10], double beta)
{
    double p[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) s = fma(m[sidx(i, j)], v[j], s);
        p[i] = beta * s;
    }
    double pv = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) pv = fma(p[j], v[j], pv);
    const double kk = 0.5 * beta * pv;
    double w[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) w[i] = (i < N) ? fma(-kk, v[i], p[i]) : p[i];
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            // m_ij -= v_i w_j + w_i v_j   (v_i = 0 for i >= N)
            double t = m[sidx(i, j)];
            if (i < N) t = fma(-v[i], w[j], t);
            if (j < N) t = fma(-w[i], v[j], t);
            m[sidx(i, j)] = t;
        }
}

// x <- (I - beta v v') x,  v with N leading non-zeros
template <int N>
CVX_HD void reflect_vec(double x[10], const double v[10], double beta)
{
    double d = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) d = fma(v[j], x[j], d);
    d *= beta;
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = fma(-d, v[j], x[j]);
}

// Cholesky of  shift I - B  (B 8x8 packed lower in the strided view Bs) and the solve
// (shift I - B) x = rhs.  Returns false when a pivot is not positive (the matrix is not positive
// definite: B has an eigenvalue >= shift).
template <int S>
CVX_HD bool chol8_solve(Arr<S> Bs, double shift, const double rhs[8], double x[8], bool want_solve)
{
    double l[36];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) l[sidx(i, j)] = ((i == j) ? shift : 0.0) - Bs[sidx(i, j)];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double d = l[sidx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-l[sidx(j, k)], l[sidx(j, k)], d);
        ok = ok && (d > 1e-300);
        const double id = cvx_rsqrt(ok ? d : 1.0);
        l[sidx(j, j)] = id;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) {
            double t = l[sidx(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-l[sidx(i, k)], l[sidx(j, k)], t);
            l[sidx(i, j)] = t * id;
        }
    }
    if (want_solve) {
        double y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double t = rhs[i];
#pragma unroll
            for (int k = 0; k < i; ++k) t = fma(-l[sidx(i, k)], y[k], t);
            y[i] = t * l[sidx(i, i)];
        }
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int i = 7 - ii;
            double t = y[i];
#pragma unroll
            for (int k = i + 1; k < 8; ++k) t = fma(-l[sidx(k, i)], x[k], t);
            x[i] = t * l[sidx(i, i)];
        }
    }
    return ok;
}

// One refinement of the tracked pairs (U: u_k[i] at U[10 k + i]; TH: th_0, th_1) for the matrix M.
// Bs: 36 doubles of per-problem scratch.  corr2 receives the largest squared correction |x_i|^2 of a
// positive slot (a large value means the step was far from converged: the caller repeats it).
// `any(pred)`: true if pred holds for this lane or for any lane executing alongside it (an
// over-approximation is fine): used so that the rarely needed third factorisation is skipped by whole warps.
template <int S, class AnyFn>
CVX_HD int track_step(Arr<S> M, Arr<S> U, Arr<S> TH, Arr<S> Bs, double& corr2, const AnyFn& any)
{
#if defined(CVX_TRK_DEBUG) && !defined(__CUDA_ARCH__)
    ++g_track_steps;
#endif
    // ---- 1. reflectors: H1 u1 = -+e9,  H2 (H1 u0) = -+e8 ------------------------------------
    double v1[10], v2[10], beta1, beta2;
    {
        double u0[10], n1 = 0.0;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            v1[i] = U[10 + i];
            u0[i] = U[i];
            n1 = fma(v1[i], v1[i], n1);
        }
        const double nr1 = sqrt(n1);
        const double a9 = fabs(v1[9]);
        v1[9] += copysign(nr1, v1[9]);
        beta1 = 1.0 / (nr1 * (nr1 + a9));     // 2 / v'v
        reflect_vec<10>(u0, v1, beta1);
        double n2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            v2[i] = u0[i];
            n2 = fma(u0[i], u0[i], n2);
        }
        v2[9] = 0.0;
        const double nr2 = sqrt(n2);
        const double a8 = fabs(v2[8]);
        v2[8] += copysign(nr2, v2[8]);
        beta2 = 1.0 / (nr2 * (nr2 + a8));
        if (!(n1 > 0.0) || !(n2 > 0.0) || !isfinite(beta1) || !isfinite(beta2)) {
            corr2 = 0.0;
            return TRK_NEED_FULL;
        }
    }
    // ---- 2. M' = H2 H1 M H1 H2 in registers; blocks ---------------------------------------
    double ct0[8], ct1[8], th0, th1, jc, js;
    {
        double m[55];
#pragma unroll
        for (int e = 0; e < 55; ++e) m[e] = M[e];
        reflect_sym<10>(m, v1, beta1);
        reflect_sym<9>(m, v2, beta2);
#pragma unroll
        for (int e = 0; e < 36; ++e) Bs[e] = m[e];   // B = M'[0:8, 0:8] (packed indices 0..35)
        // 2x2 block A on coordinates (8, 9): Ritz values and rotation
        double jt;
        jacobi_cs(m[sidx(8, 8)], m[sidx(9, 9)], m[sidx(9, 8)], jc, js, jt);
        th0 = fma(-jt, m[sidx(9, 8)], m[sidx(8, 8)]);
        th1 = fma(jt, m[sidx(9, 8)], m[sidx(9, 9)]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double c8 = m[sidx(8, i)], c9 = m[sidx(9, i)];
            ct0[i] = fma(jc, c8, -js * c9);
            ct1[i] = fma(js, c8, jc * c9);
        }
    }
    // ---- 3. solves + certificates ---------------------------------------------------------
    const bool pos0 = th0 > 0.0, pos1 = th1 > 0.0;
    const double s0 = pos0 ? th0 : 0.0, s1 = pos1 ? th1 : 0.0;
    double x0[8], x1[8];
    bool ok0 = true, ok1 = true, ok2 = true;
    // slot 0, slot 1, and (only when both slots are positive) the certificate at shift 0: one rolled loop, so
    // the factorisation exists once in the instruction stream (the loop body then runs from the instruction cache)
    const int n_fact = (pos0 && pos1) ? 3 : 2;
#pragma unroll 1
    for (int q = 0; q < n_fact; ++q) {
        double rhs[8], x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rhs[i] = (q == 0) ? ct0[i] : ct1[i];
        const bool okq = chol8_solve(Bs, q == 0 ? s0 : (q == 1 ? s1 : 0.0), rhs, x, true);
        if (q == 0) {
            ok0 = okq;
#pragma unroll
            for (int i = 0; i < 8; ++i) x0[i] = x[i];
        } else if (q == 1) {
            ok1 = okq;
#pragma unroll
            for (int i = 0; i < 8; ++i) x1[i] = x[i];
        } else {
            ok2 = okq;
        }
    }
    int rc = (ok0 && ok1 && ok2) ? TRK_OK : TRK_NEED_FULL;
#if defined(CVX_TRK_DEBUG) && !defined(__CUDA_ARCH__)
    if (rc != TRK_OK) {
        // true spectrum of M by cold Jacobi (debug only)
        double tt[55], vv[100];
        for (int e = 0; e < 55; ++e) tt[e] = M[e];
        for (int i = 0; i < 100; ++i) vv[i] = (i / 10 == i % 10);
        for (int sw = 0; sw < 12; ++sw) jacobi_sweep(Arr<1>{tt}, Arr<1>{vv});
        double ev[10];
        for (int j = 0; j < 10; ++j) ev[j] = tt[sidx(j, j)];
        for (int a = 0; a < 10; ++a) for (int b2 = a + 1; b2 < 10; ++b2) if (ev[b2] > ev[a]) { double x = ev[a]; ev[a] = ev[b2]; ev[b2] = x; }
        printf("FAIL ok %d %d %d th~ %.3e %.3e | ev %.3e %.3e %.3e %.3e\n", ok0, ok1, ok2, th0, th1, ev[0], ev[1], ev[2], ev[3]);
    }
#endif
    if (!ok0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x0[i] = 0.0;
    }
    if (!ok1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x1[i] = 0.0;
    }
    // ---- 4. Rayleigh-Ritz on span{[x0; e~0], [x1; e~1]} -------------------------------------
    double x00 = 0, x01 = 0, x11 = 0, c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x00 = fma(x0[i], x0[i], x00);
        x01 = fma(x0[i], x1[i], x01);
        x11 = fma(x1[i], x1[i], x11);
        c00 = fma(ct0[i], x0[i], c00);
        c01 = fma(ct0[i], x1[i], c01);
        c10 = fma(ct1[i], x0[i], c10);
        c11 = fma(ct1[i], x1[i], c11);
    }
    corr2 = fmax(pos0 ? x00 : 0.0, pos1 ? x11 : 0.0);
    // H_ij = s_j x_i.x_j + c~_i.x_j + th~_j delta_ij  (B x_j = s_j x_j - c~_j)
    const double h00 = fma(s0, x00, c00) + th0, h11 = fma(s1, x11, c11) + th1;
    const double h01 = 0.5 * (fma(s1, x01, c01) + fma(s0, x01, c10));
    const double g00 = 1.0 + x00, g01 = x01, g11 = 1.0 + x11;
    // G = L L', orthonormal basis q0 = w0 / l00, q1 = (w1 - l10 q0) / l11
    const double i00 = cvx_rsqrt(g00);
    const double l10 = g01 * i00;
    const double i11 = cvx_rsqrt(g11 - l10 * l10);
    const double y00 = h00 * i00, y10 = (h01 - l10 * y00) * i11, y01 = h01 * i00, y11 = (h11 - l10 * y01) * i11;
    const double hp00 = y00 * i00, hp10 = y10 * i00, hp11 = (y11 - y10 * l10 * i00) * i11;
    double rc_, rs_, rt_;
    jacobi_cs(hp00, hp11, hp10, rc_, rs_, rt_);
    const double nth0 = fma(-rt_, hp10, hp00), nth1 = fma(rt_, hp10, hp11);
    if (!isfinite(nth0) || !isfinite(nth1)) {
        return TRK_NEED_FULL;
    }
    // final vectors f0 = rc q0 - rs q1, f1 = rs q0 + rc q1 with q0 = i00 w0, q1 = i11 (w1 - l10 i00 w0)
    const double a00 = rc_ * i00 + rs_ * i11 * l10 * i00, a01 = -rs_ * i11;     // f0 = a00 w0 + a01 w1
    const double a10 = rs_ * i00 - rc_ * i11 * l10 * i00, a11 = rc_ * i11;      // f1 = a10 w0 + a11 w1
    // w_i = [x_i ; e~_i],  e~_0 = (jc, -js), e~_1 = (js, jc) on coordinates (8, 9)
    double f0[10], f1[10];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f0[i] = fma(a00, x0[i], a01 * x1[i]);
        f1[i] = fma(a10, x0[i], a11 * x1[i]);
    }
    f0[8] = fma(a00, jc, a01 * js);
    f0[9] = fma(a00, -js, a01 * jc);
    f1[8] = fma(a10, jc, a11 * js);
    f1[9] = fma(a10, -js, a11 * jc);
    // ---- 5. back through the reflectors: u = H1 H2 f ---------------------------------------
    reflect_vec<9>(f0, v2, beta2);
    reflect_vec<10>(f0, v1, beta1);
    reflect_vec<9>(f1, v2, beta2);
    reflect_vec<10>(f1, v1, beta1);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        U[i] = f0[i];
        U[10 + i] = f1[i];
    }
    TH[0] = nth0;
    TH[1] = nth1;
    return rc;
}

// Z = sum over the positive slots of th_i u_i u_i'  (registers)
template <int S>
CVX_HD void track_psd(Arr<S> U, Arr<S> TH, double z[55])
{
    const double t0 = fmax(TH[0], 0.0), t1 = fmax(TH[1], 0.0);
    double a[10], b[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        a[i] = U[i];
        b[i] = U[10 + i];
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const double ar = t0 * a[r], br = t1 * b[r];
#pragma unroll
        for (int c = 0; c <= r; ++c) z[sidx(r, c)] = fma(ar, a[c], br * b[c]);
    }
}

// From a full decomposition (V, L) to the tracked form: the two largest eigenpairs.
template <int S>
CVX_HD void track_from_full(Arr<S> V, Arr<S> L, double* u_out, double* th_out)
{
    int j0 = 0, j1 = 1;
    double l0 = -1e300, l1 = -1e300;
#pragma unroll 1
    for (int j = 0; j < 10; ++j) {
        const double l = L[j];
        if (l > l0) {
            l1 = l0; j1 = j0;
            l0 = l; j0 = j;
        } else if (l > l1) {
            l1 = l; j1 = j;
        }
    }
#pragma unroll 1
    for (int i = 0; i < 10; ++i) {
        u_out[i] = V[i * 10 + j0];
        u_out[10 + i] = V[i * 10 + j1];
    }
    th_out[0] = l0;
    th_out[1] = l1;
}

// ---------------------------------------------------------------------------------------
// Pre-pass record of the tracked solver (doubles):  Q/rho 45 | rho | M 55 | u0 10 | u1 10 | th0 th1 | it
// (rho = NaN marks a non-finite problem; the same record is the hand-over entry to the warp kernel).
// ---------------------------------------------------------------------------------------
constexpr int TR_Q = 0, TR_RHO = 45, TR_M = 46, TR_U = 101, TR_TH = 121, TR_IT = 123, TR_FLAGS = 124;
constexpr int TRK_DOUBLES = 126;
static_assert(TRK_DOUBLES <= PRE_DOUBLES, "the tracked record reuses the pre-pass record");

// Early phase (pre-pass kernel, lane-parallel): assembly has filled rec[0..45]; V receives the start
// decomposition; run Opts::early plain DR iterations with the full decomposition, then keep the two
// largest eigenpairs.  V 100, M 55, T 56, L 10: strided work arrays (shared memory).
// Returns the third-largest eigenvalue of the iterate the problem is tracked from: the closer it is to zero (or the
// more positive), the more likely a third positive eigenvalue shows up a few iterations later and the certificate fails
// (host build: every PnPL 8+4 / PnP-8 problem that was handed back was among the 2 % with the largest value, 60 % of the
// PnL-6 ones among the top 10 %).  The pre-pass uses it to put those problems at the head of the work queue.
template <int S>
CVX_HD double track_early(double* rec, const Opts& o, Arr<S> V, Arr<S> M, Arr<S> T, Arr<S> L)
{
    LaneState st;
    // start_decomposition wrote V (100) and lambda (10) into rec[PRE_V..]; problem_begin copies them in
    problem_begin(rec, o, V, M, L, GArr{rec, 1}, st);
    T[55] = 0.0;
    Opts oe = o;
    oe.anderson = false;     // no acceleration, no plateau logic in the first iterations
    if (st.finite) {
#pragma unroll 1
        for (int k = 0; k < o.early; ++k) {
            pass_dr(oe, V, M, T, L, GArr{rec, 1}, st);
            if (!st.iterating) break;      // converged / capped already (max_iters < early)
            pass_eig(oe, V, M, T, L, GArr{rec, 1}, st);
        }
    }
    double u[20], th[2];
    track_from_full(V, L, u, th);
#pragma unroll 5
    for (int e = 0; e < 55; ++e) rec[TR_M + e] = M[e];
#pragma unroll 4
    for (int e = 0; e < 20; ++e) rec[TR_U + e] = u[e];
    rec[TR_TH] = th[0];
    rec[TR_TH + 1] = th[1];
    rec[TR_IT] = (double)st.it;
    // bit 0: the DR loop is over already (converged within the early iterations or max_iters <= early);
    // bit 1: converged
    rec[TR_FLAGS] = (double)((st.iterating ? 0 : 1) | (st.converged ? 2 : 0));
    // third largest eigenvalue
    double a0 = -1e300, a1 = -1e300, a2 = -1e300;
#pragma unroll 1
    for (int j = 0; j < 10; ++j) {
        const double l = L[j];
        if (l > a0) {
            a2 = a1; a1 = a0; a0 = l;
        } else if (l > a1) {
            a2 = a1; a1 = l;
        } else if (l > a2) {
            a2 = l;
        }
    }
    return a2;
}

// ---------------------------------------------------------------------------------------
// Per-problem state machine of the tracked solver (persistent kernel, one thread per problem).
// Work arrays (strided, shared memory): M 55, G 56 (step, zero padded), U 20, TH 2, BS 36, QR 45.
// ---------------------------------------------------------------------------------------
template <int S>
CVX_HD void track_begin(const double* rec, const Opts& o, Arr<S> M, Arr<S> U, Arr<S> TH, Arr<S> QR, LaneState& st)
{
    const double rho = rec[TR_RHO];
    {
        double q[45];
#pragma unroll
        for (int e = 0; e < 45; ++e) q[e] = rec[TR_Q + e];
#pragma unroll
        for (int e = 0; e < 45; ++e) QR[e] = q[e];
    }
    {
        double m[55];
#pragma unroll
        for (int e = 0; e < 55; ++e) m[e] = rec[TR_M + e];
#pragma unroll
        for (int e = 0; e < 55; ++e) M[e] = m[e];
    }
    {
        double u[22];
#pragma unroll
        for (int e = 0; e < 22; ++e) u[e] = rec[TR_U + e];
#pragma unroll
        for (int e = 0; e < 20; ++e) U[e] = u[e];
        TH[0] = u[20];
        TH[1] = u[21];
    }
    const int fl = (int)rec[TR_FLAGS];
    st.rho = rho;
    st.dobj = 0.0;
    st.phase = 0;
    st.it = (int32_t)rec[TR_IT];
    aa_reset(st.aa);
    st.res_prev = 1e300;
    st.finite = isfinite(rho);
    st.iterating = st.finite && !(fl & 1);
    st.converged = (fl & 2) != 0;
    st.bad = 0;
    if (st.finite && !st.iterating) st.phase = 1;
}

// from the FP32 first phase (WARM record: M 55 | V 100 | lambda 10 | it; V re-orthonormalised in FP64)
template <int S>
CVX_HD void track_begin_warm(const double* rec, const double* w, const Opts& o, Arr<S> M, Arr<S> U, Arr<S> TH,
                             Arr<S> QR, LaneState& st)
{
    const double rho = rec[TR_RHO];
#pragma unroll 5
    for (int e = 0; e < 45; ++e) QR[e] = rec[TR_Q + e];
#pragma unroll 5
    for (int e = 0; e < 55; ++e) M[e] = w[e];
    double u[20], th[2];
    track_from_full(Arr<1>{const_cast<double*>(w) + 55}, Arr<1>{const_cast<double*>(w) + 155}, u, th);
#pragma unroll 4
    for (int e = 0; e < 20; ++e) U[e] = u[e];
    TH[0] = th[0];
    TH[1] = th[1];
    st.rho = rho;
    st.dobj = 0.0;
    st.phase = 0;
    st.it = (int32_t)w[165];
    aa_reset(st.aa);
    st.res_prev = 1e300;
    st.finite = isfinite(rho);
    st.iterating = st.finite && st.it < o.max_iters;
    st.converged = false;
    st.bad = 0;
    if (st.finite && !st.iterating) st.phase = 1;
}

// What follows a DR iteration with squared residual `res` (same rules as pass_dr): end of the DR loop, plateau jump
// (applied to the entries [e_lo, e_hi) of M: all of them, or one thread's share in the role-split solver), Anderson
// bookkeeping.  Returns true when the problem wants an Anderson step on the iterate it just produced.
template <int S>
CVX_HD bool track_decide(const Opts& o, LaneState& st, double res, Arr<S> M, Arr<S> G, int e_lo, int e_hi)
{
    ++st.it;
#if defined(CVX_TRACE) && !defined(__CUDA_ARCH__)
    printf("trk it %d res %.3e aa_mask %u\n", st.it, sqrt(res), st.aa.mask);
#endif
    if (!(res > o.eps2) && (st.bad == 0 || !(res <= o.eps2))) {  // also leaves on NaN; never converges on an uncertified projection
        st.converged = (res <= o.eps2);
        st.iterating = false;
        st.phase = 1;
        return false;
    }
    if (st.it >= o.max_iters) {
        st.iterating = false;
        st.phase = 1;
        return false;
    }
    if (!o.anderson) return false;
    if (st.bad > 0) {   // the step came from an uncertified projection: keep it out of the accelerator and the plateau detector
        aa_reset(st.aa);
        st.res_prev = 1e300;
        return false;
    }
    {
        const int tau = plateau_update(st.phase, res, st.res_prev);
        if (tau > 0) {
            const double ft = (double)tau;
#pragma unroll 1
            for (int e = e_lo; e < e_hi; ++e) M[e] = fma(ft, G[e], M[e]);
            aa_reset(st.aa);
            st.res_prev = res;
            return false;
        }
    }
    const bool tail = res < o.aa_on2;
    if (!tail || res > 4.0 * st.res_prev) aa_reset(st.aa);
    st.res_prev = res;
    return tail;
}

// Part 1 of a pass: one DR iteration from the tracked pairs (same rules as pass_dr).
template <int S>
CVX_HD bool track_pass_dr(const Opts& o, Arr<S> M, Arr<S> G, Arr<S> U, Arr<S> TH, Arr<S> QR, LaneState& st)
{
    if (!st.finite || !st.iterating) return false;
    ZRank2 zf;
    {
        const double t0 = fmax(TH[0], 0.0), t1 = fmax(TH[1], 0.0);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            zf.a[i] = U[i];
            zf.b[i] = U[10 + i];
            zf.ta[i] = t0 * zf.a[i];
            zf.tb[i] = t1 * zf.b[i];
        }
    }
    const double res = dr_affine_update(M, G, QR, o.alpha, 1.0 / o.sigma, o.rowk, zf);
    return track_decide(o, st, res, M, G, 0, 55);
}

// Part 2 of a pass: refine the tracked pairs for the new M; penalty rescale; end of the DR loop.
// Returns 1 when the problem is done (U, TH hold the positive eigenpairs of the final iterate and
// st.dobj the dual objective), 0 to continue, -1 when the certificate failed (hand the problem over).
#ifndef CVX_TRK_REPEAT2
#define CVX_TRK_REPEAT2 1e-4   // squared correction above which a failed certificate is re-evaluated
#endif
template <int S, class AnyFn>
CVX_HD int track_pass_eig(const Opts& o, Arr<S> M, Arr<S> U, Arr<S> TH, Arr<S> BS, Arr<S> QR, LaneState& st,
                          const AnyFn& any)
{
    if (!st.finite) return 1;
#if defined(CVX_TRK_DEBUG) && !defined(__CUDA_ARCH__)
    ++g_track_passes;
#endif
    double corr2 = 0.0;
    int rc = TRK_OK;
    // One step per DR iteration is enough even right after a plateau jump or an extrapolation (measured,
    // host build: repeating the step whenever the correction is large changes neither the iteration
    // counts nor the poses).  Only a FAILED certificate is looked at again (up to three times) when the
    // step it was computed in was far from converged: B is the complement of the vectors BEFORE the step,
    // and with a large correction it still contains what the slots have not picked up yet.
    // (One rolled loop so that the step exists once in the instruction stream.)
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        rc = track_step(M, U, TH, BS, corr2, any);
        if (!(rc != TRK_OK && corr2 > CVX_TRK_REPEAT2)) break;
    }
#ifndef CVX_TRK_TOLERATE
#define CVX_TRK_TOLERATE 8   // hand-backs, host build, 1000 problems each, 0 / 4 / 8 / 16: PnP-8 1.2 / 0.7 / 0.2 / 0.1 %, PnL-6 6.8 / 5.2 / 4.1 / 3.8 %
#endif
    if (rc != TRK_OK) {
        // A third positive eigenvalue is usually a transient of the first 10-20 iterations (host build: 92 % of the
        // failed certificates; median iteration 11).  Any M is a valid DR state, so the problem may go on with the
        // two tracked pairs -- an inexact projection -- for up to CVX_TRK_TOLERATE iterations, out of the accelerator's
        // sight and unable to converge, before it is handed back.
        if (!st.iterating || st.bad >= CVX_TRK_TOLERATE || !isfinite(corr2)) return -1;
        ++st.bad;
    } else {
        st.bad = 0;
    }
    if (st.iterating) {
        const double rf = rescale_factor(st.it);
        if (rf > 0.0) {
            // M = Z + N (N the negative part): N <- N / rf, i.e. M <- Z (1 - 1/rf) + M / rf; Q/rho likewise
            const double ic = 1.0 / rf;
            double z[55];
            track_psd(U, TH, z);
#pragma unroll
            for (int e = 0; e < 55; ++e) M[e] = fma(ic, (double)M[e] - z[e], z[e]);
#pragma unroll 1
            for (int e = 0; e < 45; ++e) QR[e] = QR[e] * ic;
            st.rho *= rf;
            aa_reset(st.aa);
            st.res_prev = 1e300;
        }
        return 0;
    }
    // DR loop over: polish until the correction vanishes
    if (corr2 > 1e-28 && st.phase < 8) {
        st.phase += 2;      // (bits 1.. count the polishing passes; at most 3)
        return 0;
    }
    // dual objective (see dual_objective): U_neg = Z - M restricted to what the feasible point needs
    {
        const double t0 = fmax(TH[0], 0.0), t1 = fmax(TH[1], 0.0);
        double trz = 0.0, tq = 0.0, trm = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            trz = fma(t0 * U[i], U[i], fma(t1 * U[10 + i], U[10 + i], trz));
            trm += M[sidx(i, i)];
            tq += QR[sidx(i, i)];
        }
        const double z99 = fma(t0 * U[9], U[9], t1 * U[19] * U[19]);
        const double tr9 = trm - trz, u99 = M[sidx(9, 9)] - z99;
        st.dobj = st.rho * ((tq + tr9) * (1.0 / 3.0) + o.sigma * o.sigma * u99);
    }
    return 1;
}

// Park the result in the format finish_kernel reads (V 100 | lambda 10 | dobj | status): the
// eigen-decomposition of the UNSCALED Z = D^-1 Z' D^-1 restricted to its (at most two) positive
// eigenpairs -- what the reference thresholds at 1e-3 (cvxpnpl.py:499-502); all other columns are zero
// with eigenvalue -1.
template <int S>
CVX_HD void track_park(const Opts& o, Arr<S> U, Arr<S> TH, const LaneState& st, double* park, int32_t* iters_out)
{
    int32_t status = ST_NAN;
    const double isig = 1.0 / o.sigma;
    double t0 = TH[0], t1 = TH[1];
    if (st.finite) {
        status = st.converged ? ST_OK : ST_MAX_ITERS;
        if (!isfinite(t0) || !isfinite(t1)) status = ST_NAN;
    }
    double a[10], b[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        a[i] = U[i];
        b[i] = U[10 + i];
    }
    a[9] *= isig;      // D^-1 u
    b[9] *= isig;
    double la = t0, lb = t1;
    if (t0 > 1e-3 && t1 > 1e-3) {
        // rank 2 in the scaled problem: eigenpairs of Z = [sqrt(t0) a, sqrt(t1) b] [..]' from its 2x2 Gram matrix
        const double r0 = sqrt(t0), r1 = sqrt(t1);
        double gaa = 0, gab = 0, gbb = 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            a[i] *= r0;
            b[i] *= r1;
            gaa = fma(a[i], a[i], gaa);
            gab = fma(a[i], b[i], gab);
            gbb = fma(b[i], b[i], gbb);
        }
        double c, s, t;
        jacobi_cs(gaa, gbb, gab, c, s, t);
        la = fma(-t, gab, gaa);
        lb = fma(t, gab, gbb);
        const double ia = 1.0 / sqrt(la), ib = 1.0 / sqrt(lb);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const double ai = a[i], bi = b[i];
            a[i] = fma(c, ai, -s * bi) * ia;
            b[i] = fma(s, ai, c * bi) * ib;
        }
    } else {
        // rank <= 1 for certain (eigenvalues of Z lie in [lam'/sigma^2, lam']): D^-1 u is all the rank-1
        // extraction needs (it divides by the homogeneous component), the eigenvalue stays the scaled one
    }
#pragma unroll 1
    for (int i = 0; i < 10; ++i) {
        park[i * 10] = a[i];
        park[i * 10 + 1] = b[i];
#pragma unroll
        for (int j = 2; j < 10; ++j) park[i * 10 + j] = 0.0;
    }
    park[100] = la;
    park[101] = lb;
#pragma unroll
    for (int j = 2; j < 10; ++j) park[100 + j] = -1.0;
    park[110] = (status != ST_NAN) ? st.dobj : nan("");
    park[111] = (double)status;
    *iters_out = st.it;
}

// Hand the problem to the warp-per-problem kernel: its record (Q/rho, rho, M, iteration count) goes back
// into the pre-pass entry it came from.
template <int S>
CVX_HD void track_handoff(Arr<S> M, Arr<S> QR, const LaneState& st, double* rec)
{
#pragma unroll 5
    for (int e = 0; e < 45; ++e) rec[TR_Q + e] = QR[e];
    rec[TR_RHO] = st.rho;
#pragma unroll 5
    for (int e = 0; e < 55; ++e) rec[TR_M + e] = M[e];
    rec[TR_IT] = (double)st.it;
    rec[TR_FLAGS] = (double)((st.iterating ? 0 : 1) | (st.converged ? 2 : 0));
}

struct AnyLane {   // host build / lane-parallel stage kernels: every "warp" is one lane
    CVX_HD bool operator()(bool f) const { return f; }
};

}  // namespace cvx
