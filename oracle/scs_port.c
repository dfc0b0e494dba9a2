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
 * coordinates.  SCS's data equilibration and Anderson acceleration are NOT
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
#define M_ 77
#define NZ 22
#define PS 10

typedef struct {
    double A[M_ * N_];      /* dense row-major */
    double b[M_];
    double L[N_ * N_];      /* Cholesky factor of I + A^T A (lower) */
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
int scs_port_setup(const double *A, const double *b)
{
    memcpy(W.A, A, sizeof(W.A));
    memcpy(W.b, b, sizeof(W.b));
    for (int i = 0; i < N_; ++i)
        for (int j = 0; j < N_; ++j) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < M_; ++k) s += A[k * N_ + i] * A[k * N_ + j];
            W.L[i * N_ + j] = s;
        }
    chol55(W.L);
    int nz = 0;
    for (int k = 0; k < M_; ++k) {
        W.rp[k] = nz;
        for (int j = 0; j < N_; ++j)
            if (A[k * N_ + j] != 0.0) { W.ci[nz] = j; W.cv[nz] = A[k * N_ + j]; ++nz; }
    }
    W.rp[M_] = nz;
    W.ready = 1;
    return 0;
}

static void Amul(const double *x, double *y)      /* y = A x */
{
    for (int k = 0; k < M_; ++k) {
        double s = 0;
        for (int e = W.rp[k]; e < W.rp[k + 1]; ++e) s += W.cv[e] * x[W.ci[e]];
        y[k] = s;
    }
}
static void ATmul(const double *y, double *x)     /* x = A^T y */
{
    for (int j = 0; j < N_; ++j) x[j] = 0;
    for (int k = 0; k < M_; ++k) {
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
    for (int k = 0; k < M_; ++k) y[k] += wy[k];
}

/* cyclic Jacobi symmetric eigensolver, cold start, PS x PS */
static void jacobi_eig(double *S, double *V)
{
    for (int i = 0; i < PS; ++i)
        for (int j = 0; j < PS; ++j) V[i * PS + j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dg = 0;
        for (int i = 0; i < PS; ++i)
            for (int j = 0; j < PS; ++j)
                if (i != j) off += S[i * PS + j] * S[i * PS + j];
                else dg += S[i * PS + j] * S[i * PS + j];
        if (off <= 1e-34 * (dg + off) || off == 0.0) break;
        for (int p = 0; p < PS - 1; ++p)
            for (int q = p + 1; q < PS; ++q) {
                double apq = S[p * PS + q];
                /* negligible pivot: skip (also keeps denormals out of the loop) */
                if (fabs(apq) <= 1e-20 * (fabs(S[p * PS + p]) + fabs(S[q * PS + q])) || fabs(apq) < 1e-290) continue;
                double theta = (S[q * PS + q] - S[p * PS + p]) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < PS; ++k) {
                    double akp = S[k * PS + p], akq = S[k * PS + q];
                    S[k * PS + p] = c * akp - s * akq;
                    S[k * PS + q] = s * akp + c * akq;
                }
                for (int k = 0; k < PS; ++k) {
                    double apk = S[p * PS + k], aqk = S[q * PS + k];
                    S[p * PS + k] = c * apk - s * aqk;
                    S[q * PS + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < PS; ++k) {
                    double vkp = V[k * PS + p], vkq = V[k * PS + q];
                    V[k * PS + p] = c * vkp - s * vkq;
                    V[k * PS + q] = s * vkp + c * vkq;
                }
            }
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
    jacobi_eig(S, V);
    double lam[PS];
    for (int i = 0; i < PS; ++i) lam[i] = S[i * PS + i] > 0 ? S[i * PS + i] : 0.0;
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
    const int n = N_, m = M_, l = N_ + M_ + 1;
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
        proj_psd_svec(u + n + NZ);
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
