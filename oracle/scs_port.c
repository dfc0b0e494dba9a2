/*
 * oracle/scs_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the conic solve that the reference delegates to the
 * third-party package `scs` (reference: cvxpnpl.py:485-489, requirements.txt:4
 * `scs>=2.0.0`, un-pinned, NOT vendored and NOT installable here).  Because the
 * SCS sources are absent, this file restates the *published* algorithm
 * (O'Donoghue, Chu, Parikh, Boyd, "Conic Optimization via Operator Splitting
 * and Homogeneous Self-Dual Embedding", JOTA 2016): Douglas-Rachford splitting
 * on the homogeneous self-dual embedding with over-relaxation, a cached
 * factorisation of the (data independent) linear system, and Euclidean
 * projection onto  {0}^22 x S_+^10  in SCS's svec (sqrt(2) off-diagonal)
 * coordinates (eigen-decomposition by Householder tridiagonalisation + implicit
 * QL, the dsyev algorithm SCS obtains from LAPACK).  SCS's data equilibration and Anderson acceleration are NOT
 * restated: they change the trajectory, not the fixed point.
 *
 * PARITY UNPINNED at the SCS boundary: the reference holds no golden vectors
 * for scs.solve (it has no tests at all).  What pins this
 * file: the three known-answer examples (examples/pnp.py, pnl.py, pnpl.py) and
 * the solver-independent KKT certificate in oracle/kkt.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this.  The product path never does.
 *
 * Problem solved (the shape cvxpnpl.py:387-448 builds, generalised only in c):
 *     minimise  c^T x   s.t.  A x + s = b,  s in {0}^22 x S_+^10,
 * with A 77x55 given dense row-major by the caller (the caller passes the
 * reference's own _A so nothing about the constraint set is hard-coded here).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N_ 55
#define M_ 77   /* maximum number of rows: 22 equalities + 55 cone rows */
#define PS 10

typedef struct {
    double A[M_ * N_];      /* dense row-major */
    double b[M_];
    double L[N_ * N_];      /* Cholesky factor of I + A^T A (lower) */
    int m;                  /* rows in use: nz + 55 */
    int nz;                 /* equality (zero cone) rows: 22, or 16 for the "rc" ablation */
    int rp[M_ + 1];         /* CSR of A (the reference ships A sparse: csc_matrix, cvxpnpl.py:442) */
    int ci[M_ * N_];
    double cv[M_ * N_];
    int ready;
} scs_port_work;

static scs_port_work W;

static void chol55(double *G)
{
    for (int j = 0; j < N_; ++j) {
        double d = G[j * N_ + j];
        for (int k = 0; k < j; ++k) d -= G[j * N_ + k] * G[j * N_ + k];
        d = sqrt(d);
        G[j * N_ + j] = d;
        for (int i = j + 1; i < N_; ++i) {
            double s = G[i * N_ + j];
            for (int k = 0; k < j; ++k) s -= G[i * N_ + k] * G[j * N_ + k];
            G[i * N_ + j] = s / d;
        }
    }
}

static void chol_solve55(const double *L, double *x)
{
    for (int i = 0; i < N_; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[i * N_ + k] * x[k];
        x[i] = s / L[i * N_ + i];
    }
    for (int i = N_ - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < N_; ++k) s -= L[k * N_ + i] * x[k];
        x[i] = s / L[i * N_ + i];
    }
}

/* one-time setup: stores A, b and factors I + A^T A */
int scs_port_setup(const double *A, const double *b, int nz)
{
    if (nz < 1 || nz + 55 > M_) return -1;
    W.nz = nz;
    W.m = nz + 55;
    memset(W.A, 0, sizeof(W.A));
    memset(W.b, 0, sizeof(W.b));
    memcpy(W.A, A, sizeof(double) * W.m * N_);
    memcpy(W.b, b, sizeof(double) * W.m);
    for (int i = 0; i < N_; ++i)
        for (int j = 0; j < N_; ++j) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < W.m; ++k) s += A[k * N_ + i] * A[k * N_ + j];
            W.L[i * N_ + j] = s;
        }
    chol55(W.L);
    int nnz = 0;
    for (int k = 0; k < W.m; ++k) {
        W.rp[k] = nnz;
        for (int j = 0; j < N_; ++j)
            if (A[k * N_ + j] != 0.0) { W.ci[nnz] = j; W.cv[nnz] = A[k * N_ + j]; ++nnz; }
    }
    W.rp[W.m] = nnz;
    W.ready = 1;
    return 0;
}

static void Amul(const double *x, double *y)      /* y = A x */
{
    for (int k = 0; k < W.m; ++k) {
        double s = 0;
        for (int e = W.rp[k]; e < W.rp[k + 1]; ++e) s += W.cv[e] * x[W.ci[e]];
        y[k] = s;
    }
}
static void ATmul(const double *y, double *x)     /* x = A^T y */
{
    for (int j = 0; j < N_; ++j) x[j] = 0;
    for (int k = 0; k < W.m; ++k) {
        const double yk = y[k];
        for (int e = W.rp[k]; e < W.rp[k + 1]; ++e) x[W.ci[e]] += W.cv[e] * yk;
    }
}

/* applies the inverse of  M = [[I, A^T], [-A, I]]  (paper section 4.1):
 *   x = (I + A^T A)^{-1} (wx - A^T wy),   y = wy + A x                      */
static void solveM(const double *wx, const double *wy, double *x, double *y)
{
    double t[N_];
    ATmul(wy, t);
    for (int j = 0; j < N_; ++j) x[j] = wx[j] - t[j];
    chol_solve55(W.L, x);
    Amul(x, y);
    for (int k = 0; k < W.m; ++k) y[k] += wy[k];
}

/* Symmetric eigen-decomposition, PS x PS: Householder tridiagonalisation followed
 * by implicit-shift QL (the algorithm behind LAPACK dsteqr/dsyev, which SCS calls
 * for its PSD projections).  On exit the columns of Z are the eigenvectors and
 * d[] the eigenvalues.  Z holds the symmetric input on entry. */
static void sym_eig(double *Z, double *d)
{
    const int n = PS;
    double e[PS];
    /* --- Householder reduction to tridiagonal form (accumulating transforms) --- */
    for (int i = n - 1; i > 0; --i) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += fabs(Z[i * n + k]);
            if (scale == 0.0) {
                e[i] = Z[i * n + l];
            } else {
                for (int k = 0; k <= l; ++k) {
                    Z[i * n + k] /= scale;
                    h += Z[i * n + k] * Z[i * n + k];
                }
                double f = Z[i * n + l];
                double g = (f >= 0.0) ? -sqrt(h) : sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                Z[i * n + l] = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    Z[j * n + i] = Z[i * n + j] / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += Z[j * n + k] * Z[i * n + k];
                    for (int k = j + 1; k <= l; ++k) g += Z[k * n + j] * Z[i * n + k];
                    e[j] = g / h;
                    f += e[j] * Z[i * n + j];
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = Z[i * n + j];
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) Z[j * n + k] -= (f * e[k] + g * Z[i * n + k]);
                }
            }
        } else {
            e[i] = Z[i * n + l];
        }
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += Z[i * n + k] * Z[k * n + j];
                for (int k = 0; k <= l; ++k) Z[k * n + j] -= g * Z[k * n + i];
            }
        }
        d[i] = Z[i * n + i];
        Z[i * n + i] = 1.0;
        for (int j = 0; j <= l; ++j) Z[j * n + i] = Z[i * n + j] = 0.0;
    }
    /* --- implicit QL on the tridiagonal matrix ---------------------------------- */
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) break;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = Z[k * n + i + 1];
                        Z[k * n + i + 1] = s * Z[k * n + i] + c * f;
                        Z[k * n + i] = c * Z[k * n + i] - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
}

/* projection of svec-coordinates (column-major lower triangle, off-diag * sqrt2)
 * onto the PSD cone */
static void proj_psd_svec(double *v)
{
    double S[PS * PS], V[PS * PS];
    const double is2 = 0.70710678118654752440, s2 = 1.41421356237309504880;
    int k = 0;
    for (int j = 0; j < PS; ++j)
        for (int i = j; i < PS; ++i, ++k) {
            double e = (i == j) ? v[k] : v[k] * is2;
            S[i * PS + j] = e;
            S[j * PS + i] = e;
        }
    double lam[PS];
    memcpy(V, S, sizeof(V));
    sym_eig(V, lam);
    for (int i = 0; i < PS; ++i) lam[i] = lam[i] > 0 ? lam[i] : 0.0;
    k = 0;
    for (int j = 0; j < PS; ++j)
        for (int i = j; i < PS; ++i, ++k) {
            double e = 0;
            for (int l = 0; l < PS; ++l) e += lam[l] * V[i * PS + l] * V[j * PS + l];
            v[k] = (i == j) ? e : e * s2;
        }
}

static double ninf(const double *x, int n)
{
    double m = 0;
    for (int i = 0; i < n; ++i) { double a = fabs(x[i]); if (a > m) m = a; }
    return m;
}
static double dot(const double *a, const double *b, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/*
 * Solve for one cost vector c.  Outputs x[55], y[77], s[77],
 * info[8] = {pobj, dobj, res_pri, res_dual, gap, iters, status, tau}
 * status: 1 solved, 2 solved_inaccurate (hit max_iters), -1 infeasible/unbounded
 * (x filled with NaN in that case, mirroring SCS).
 */
int scs_port_solve(const double *c_in, double eps_abs, double eps_rel, int max_iters,
                   double alpha, double cscale, double *x_out, double *y_out, double *s_out,
                   double *info)
{
    if (!W.ready) return -100;
    const int n = N_, m = W.m, l = N_ + W.m + 1;
    double u[N_ + M_ + 1], v[N_ + M_ + 1], ut[N_ + M_ + 1], w[N_ + M_ + 1];
    double hx[N_], hy[M_], gx[N_], gy[M_];   /* h = (c, b);  g = M^{-1} h */
    double x[N_], y[M_], s[M_], Ax[M_], ATy[N_];

    /* cost scaling: stands in for SCS's data normalisation / `scale` parameter.
     * It rescales the dual variables (y -> cscale * y) and leaves x, and hence
     * the fixed point in x, unchanged.  Residuals below are reported unscaled. */
    double c[N_];
    for (int j = 0; j < n; ++j) c[j] = cscale * c_in[j];
    memcpy(hx, c, sizeof(hx));
    memcpy(hy, W.b, sizeof(hy));
    solveM(hx, hy, gx, gy);                   /* g = M^{-1} h */
    double hTg = dot(hx, gx, n) + dot(hy, gy, m);

    for (int i = 0; i < l; ++i) { u[i] = 0; v[i] = 0; }
    u[l - 1] = sqrt((double)l);
    v[l - 1] = sqrt((double)l);

    int it, status = 2;
    double pobj = NAN, dobj = NAN, rp = NAN, rd = NAN, gap = NAN, tau = NAN;
    const double nb = ninf(W.b, m), nc = ninf(c_in, n);
    for (it = 0; it < max_iters; ++it) {
        /* --- subspace projection: ut = (I + Q)^{-1} (u + v) --------------- */
        for (int i = 0; i < l; ++i) w[i] = u[i] + v[i];
        double wt = w[l - 1];
        /* M ut_xy = w_xy - ut_tau h ;  ut_tau = w_tau + h^T ut_xy
         * => ut_tau = (w_tau + h^T M^{-1} w_xy) / (1 + h^T M^{-1} h)           */
        double zx[N_], zy[M_];
        solveM(w, w + n, zx, zy);
        double utau = (wt + dot(hx, zx, n) + dot(hy, zy, m)) / (1.0 + hTg);
        for (int j = 0; j < n; ++j) ut[j] = zx[j] - utau * gx[j];
        for (int k = 0; k < m; ++k) ut[n + k] = zy[k] - utau * gy[k];
        ut[l - 1] = utau;

        /* --- cone projection with over-relaxation ------------------------ */
        for (int j = 0; j < n; ++j) {           /* x block: free (not relaxed, as SCS) */
            u[j] = ut[j] - v[j];
        }
        double rel[M_ + 1];
        for (int k = 0; k <= m; ++k) rel[k] = alpha * ut[n + k] + (1 - alpha) * u[n + k];
        for (int k = 0; k <= m; ++k) u[n + k] = rel[k] - v[n + k];
        /* y block: dual cone = R^22 x S_+^10 ; tau >= 0 */
        proj_psd_svec(u + n + W.nz);
        if (u[l - 1] < 0) u[l - 1] = 0;
        /* --- dual update -------------------------------------------------- */
        for (int j = 0; j < n; ++j) v[j] = 0.0;  /* r = 0 for the free block */
        for (int k = 0; k <= m; ++k) v[n + k] += u[n + k] - rel[k];

        /* --- termination (every 10 iterations) --------------------------- */
        if ((it % 10) == 9 || it == max_iters - 1) {
            tau = u[l - 1];
            if (tau > 1e-12) {
                for (int j = 0; j < n; ++j) x[j] = u[j] / tau;
                for (int k = 0; k < m; ++k) { y[k] = u[n + k] / (tau * cscale); s[k] = v[n + k] / tau; }
                Amul(x, Ax);
                ATmul(y, ATy);
                double t1[M_], t2[N_];
                for (int k = 0; k < m; ++k) t1[k] = Ax[k] + s[k] - W.b[k];
                for (int j = 0; j < n; ++j) t2[j] = ATy[j] + c_in[j];
                rp = ninf(t1, m);
                rd = ninf(t2, n);
                pobj = dot(c_in, x, n);
                dobj = -dot(W.b, y, m);
                gap = fabs(pobj - dobj);
                double np_ = fmax(fmax(ninf(Ax, m), ninf(s, m)), nb);
                double nd_ = fmax(ninf(ATy, n), nc);
                double ng_ = fmax(fabs(pobj), fabs(dobj));
                if (rp <= eps_abs + eps_rel * np_ && rd <= eps_abs + eps_rel * nd_ &&
                    gap <= eps_abs + eps_rel * ng_) {
                    status = 1;
                    ++it;
                    break;
                }
            }
        }
    }
    tau = u[l - 1];
    if (!(tau > 1e-12)) {
        for (int j = 0; j < n; ++j) x_out[j] = NAN;
        for (int k = 0; k < m; ++k) { y_out[k] = NAN; s_out[k] = NAN; }
        status = -1;
    } else {
        for (int j = 0; j < n; ++j) x_out[j] = u[j] / tau;
        for (int k = 0; k < m; ++k) { y_out[k] = u[n + k] / (tau * cscale); s_out[k] = v[n + k] / tau; }
    }
    info[0] = pobj; info[1] = dobj; info[2] = rp; info[3] = rd; info[4] = gap;
    info[5] = (double)it; info[6] = (double)status; info[7] = tau;
    return status;
}
