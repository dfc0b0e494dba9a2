// pnpl_solve.cuh -- the per-problem driver: assembly -> DR/ADMM on the 10x10 SDP
// -> pose extraction.  Shared by the fused CUDA kernel and (for debugging only)
// the host harness under tests/.
#pragma once

#include "pnpl_core.cuh"
#include "pnpl_extract.cuh"

namespace cvx {

struct Opts {
    double eps2;      // squared fixed-point tolerance
    double alpha;     // over-relaxation
    double rho_rel;   // penalty relative to ||Q||_F
    int max_iters;
    int sweeps;       // Jacobi sweeps per iteration (warm started)
};

struct Problem {
    const double* K;
    const double* pts_2d;
    const double* pts_3d;
    const double* line_2d;
    const double* line_3d;
    int n_pts, n_lines;
};

struct Result {
    int32_t n_poses, status, iters;
    double pobj, dobj;
};

// ---------------------------------------------------------------------------------
// DR/ADMM loop for one problem.  On entry qr holds Q/rho.  On exit V, lam hold a
// converged eigen-decomposition of the final DR iterate M (Z = V max(lam,0) V').
// ---------------------------------------------------------------------------------
template <int S, class QR>
CVX_HD int dr_solve(Arr<S> V, Arr<S> M, QR qr, const Opts& o, double lam[10], bool& converged)
{
    // start: Z0 = blkdiag(I/3, 1) (feasible for the diagonal block), U0 = 0
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#pragma unroll
        for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) M[sidx(i, j)] = (i == j) ? (i == 9 ? 1.0 : 1.0 / 3.0) : 0.0;
        lam[i] = (i == 9) ? 1.0 : 1.0 / 3.0;
    }
    converged = false;
    int it = 0;
    bool iterating = true;
    // One loop body serves both the DR iterations (one warm-started sweep each) and
    // the final passes that drive the eigen-decomposition of the last iterate to
    // full convergence, so the (large, unrolled) sweep code exists once.
    for (int guard = 0; guard < o.max_iters + 40; ++guard) {
        if (iterating) {
            double z[55];
            const double res = dr_step(M, V, lam, qr, o.alpha, z);
            ++it;
            if (!(res > o.eps2)) {  // also leaves on NaN
                converged = (res <= o.eps2);
                iterating = false;
            } else if (it >= o.max_iters) {
                iterating = false;
            }
        }
        double t[55];
        rotate_into_basis_reg(M, V, t);
        double off = 0.0, dg = 0.0;
        for (int s = 0; s < (iterating ? o.sweeps : 1); ++s) off = jacobi_sweep_reg(t, V);
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            lam[j] = t[sidx(j, j)];
            dg = fma(lam[j], lam[j], dg);
        }
        if (!iterating && !(off > 1e-22 * dg)) break;
    }
    return it;
}

// Dual objective.  At the fixed point Q + rho U = sum_k y_k P_k with U = V min(lam,0) V'
// the scaled dual slack, so for ANY affine-feasible point Zf the dual objective
// y_0 = sum_k y_k <P_k, Zf> = <Q + rho U, Zf>.  Zf = blkdiag(I/3, 1) is used.
template <int S, class QR>
CVX_HD double dual_objective(Arr<S> V, const double lam[10], QR qr, double rho)
{
    double tr9 = 0, u99 = 0, tq = 0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const double ln = fmin(lam[j], 0.0);
        double s = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) s = fma(V[i * 10 + j], V[i * 10 + j], s);
        tr9 = fma(ln, s, tr9);
        u99 = fma(ln, V[90 + j] * V[90 + j], u99);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) tq += qr[sidx(i, i)];
    return rho * ((tq + tr9) * (1.0 / 3.0) + u99);
}

template <int S>
CVX_HD void write_Z(Arr<S> V, const double lam[10], bool is_nan, double* Zo)
{
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c <= r; ++c) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < 10; ++j) s = fma(fmax(lam[j], 0.0) * V[r * 10 + j], V[c * 10 + j], s);
            if (is_nan) s = nan("");
            Zo[r * 10 + c] = s;
            Zo[c * 10 + r] = s;
        }
}

// Whole path for one problem.  V (100), M (55) and QR (45) are the problem's strided
// work arrays (200 doubles per problem, all in shared memory in the CUDA kernel).
template <int S>
CVX_HD void solve_problem(const Problem& pr, const Opts& o, Arr<S> V, Arr<S> M, Arr<S> QR,
                          double* R_out, double* t_out, double* Z_out, Result& rs)
{
    // ---- assembly: Q (45) and B (27) land in the V region, which is free until the
    // DR loop initialises it; Q/rho goes to its own region ---------------------------
    double rho;
    bool finite;
    {
        Arr<S> Qs = V;
        Arr<S> Bs = V.sub(45);
        finite = assemble(pr.K, pr.pts_2d, pr.pts_3d, pr.n_pts, pr.line_2d, pr.line_3d, pr.n_lines, Qs, Bs);
        double nq = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double q = Qs[sidx(i, j)];
                nq = fma((i == j) ? 1.0 : 2.0, q * q, nq);
            }
        rho = o.rho_rel * sqrt(nq);
        finite = finite && (rho > 0.0) && isfinite(rho);
        const double ir = 1.0 / rho;
#pragma unroll
        for (int e = 0; e < 45; ++e) QR[e] = Qs[e] * ir;
    }

    int32_t status = ST_NAN;
    int it = 0;
    double lam[10];
    bool converged = false;
    if (finite) {
        it = dr_solve(V, M, QR, o, lam, converged);
        status = converged ? ST_OK : ST_MAX_ITERS;
#pragma unroll
        for (int j = 0; j < 10; ++j)
            if (!isfinite(lam[j])) status = ST_NAN;
    }
    const double dobj = (status != ST_NAN) ? dual_objective(V, lam, QR, rho) : nan("");
    if (Z_out) write_Z(V, lam, status == ST_NAN, Z_out);

    // ---- extraction: Q and B are re-assembled (cheap) into the now free M / QR regions
    Arr<S> Qs = M;
    Arr<S> Bs = QR;
    if (status != ST_NAN) assemble(pr.K, pr.pts_2d, pr.pts_3d, pr.n_pts, pr.line_2d, pr.line_3d, pr.n_lines, Qs, Bs);
    double pobj;
    rs.n_poses = extract_poses(V, lam, Qs, Bs, status, dobj, sqrt(o.eps2), R_out, t_out, pobj);
    rs.status = status;
    rs.iters = it;
    rs.pobj = pobj;
    rs.dobj = dobj;
}

}  // namespace cvx
