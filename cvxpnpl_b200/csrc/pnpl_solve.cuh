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
    double sigma;     // homogeneous-coordinate scaling (see dr_step)
    bool anderson;    // Anderson acceleration of the DR iteration (see aa_step)
    double aa_on2;    // squared residual below which Anderson acceleration is active
    double rowk;      // 1: reference SDP (22 equalities); 0: "rc" ablation (16 equalities)
    double kappa;     // dual guess of the start point: U0 = kappa Q/rho (see start_decomposition)
    int early;        // tracked solver: DR iterations run with the full decomposition in the pre-pass (pnpl_track.cuh)
};

// Default DR parameters (used where the descriptor leaves them 0), measured on seeded
// batches with the Anderson accelerator, the dual-guess start and the plateau jump in place
// (host build, 3000 problems, mean / p99 iterations).  From 8 points up rho = 0.017 ||Q||_F,
// sigma = 1.25: PnPL 8+4 52.4 / 74, PnP-8 53.8 / 86 (the landscape is flat: 0.014 ... 0.02 x
// 1.15 ... 1.4 are all within 1 %; the earlier default 0.01 / 1.5 gives 54.5 / 75 and 55.6 / 84).
// With fewer points -- lines only, minimal and near-minimal sets -- the iteration is more robust
// with rho = 0.007 ||Q||_F, sigma = 1.25 (PnL-6 111 / 587 -> 89 / 237, 4 points 474 -> 360 and
// 37 -> 24 of 400 at the cap, 6 points 82 / 598 -> 74 / 293; measured before the plateau jump).
// Over-relaxation 1.5 is as good or better than 1.3, 1.65 and 1.8 everywhere.
CVX_HD void default_params(int n_pts, double& rho_rel, double& alpha, double& sigma)
{
    const bool many = n_pts >= 8;
    if (!(rho_rel > 0)) rho_rel = many ? 0.017 : 0.007;
    if (!(sigma > 0)) sigma = 1.25;
    if (!(alpha > 0)) alpha = 1.5;
}

// Slow problems get a smaller penalty, once.  A problem still iterating after RESCALE_AT
// iterations continues with rho <- 0.35 rho: the scaled dual U = Y / rho grows by
// 1 / 0.35, i.e. in the eigenbasis of M = Z - U the negative eigenvalues are
// divided by 0.35, and so is Q / rho.  Same SDP, same fixed point; measured on
// seeded batches (mean iterations / problems at the 2500 cap): PnL-6 106 / 19 -> 95 / 5 of
// 6000, 4 points 356 / 331 -> 262 / 165, PnP-8 max 696 -> 443; well-posed problems never get
// this far (PnPL 8+4: p99.9 = 107).  Slow problems are the ones whose optimal face is nearly
// flat (two small eigenvalues of Q); there a smaller rho lets the objective pull harder.
// Problems that are still not done at 400 and at 800 iterations get another factor 0.5 each
// (PnL-6: 30 -> 7 of 20000 at the cap; 4 points: 121 -> 64 of 3000).
#ifndef CVX_RESCALE_AT
#define CVX_RESCALE_AT 150
#endif
constexpr int RESCALE_AT = CVX_RESCALE_AT;
constexpr int PLAT_JUMPED = 1 << 30, PLAT_OFF = 0xffff;
#ifndef CVX_PLAT_N
#define CVX_PLAT_N 5
#endif
#ifndef CVX_PLAT_TOL
#define CVX_PLAT_TOL 3e-3
#endif
#ifndef CVX_PLAT_TAU0
#define CVX_PLAT_TAU0 8
#endif
#ifndef CVX_PLAT_TAU_MAX
#define CVX_PLAT_TAU_MAX 256
#endif
CVX_HD double rescale_factor(int it) { return it == RESCALE_AT ? 0.35 : ((it == 400 || it == 800) ? 0.5 : 0.0); }

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
// ---------------------------------------------------------------------------------
// Plateau jump.  Slow problems spend hundreds of iterations walking at CONSTANT step: between two
// changes of the active face DR is (nearly) a translation M <- M + g -- the residual |g| stays the
// same to 4+ digits while the dual walks towards a far-away face -- and Anderson extrapolation has
// nothing to work with (dG = 0); a smaller penalty does not help either (the walk has the same
// length in units of rho g).  Any M is a valid DR state, so after PLAT_N iterations with the same
// residual the problem jumps tau steps ahead along g.  tau is a trust region: the first iterate
// after a jump shows how hard the jump kicked the residual (the translation is only first-order
// exact); a kick below 1.5x lets tau double for the next jump (8, 16, ... PLAT_TAU_MAX), above 1.5x
// tau stays, above 4x it is halved, and a problem whose tau shrinks to nothing stops jumping.
// Measured with the host build of these routines on seeded batches (iterations mean / p99 / max):
// PnPL 8+4 55.8 / 82 / 244 -> 55.5 / 76 / 119, the 69 slowest of 3e5 PnPL problems 252 / - / 799 ->
// 121 / - / 254, PnP-8 58.4 / 118 / 1143 -> 56.8 / 89 / 311, PnL-6 95.3 / 501 / 2500 -> 78 / 237 / 708;
// same poses (<= 3e-9 rad).
// `plat` packs the state: bits 0-1 zero (LaneState::phase lives there), bits 2-9 flat-residual streak,
// bits 10-25 tau, bit 30 "jumped last time".
// Returns the number of steps to jump now (0: none); res, res_prev are squared residuals.
// ---------------------------------------------------------------------------------
CVX_HD int plateau_update(int32_t& plat, double res, double res_prev)
{
    int tau = (plat >> 10) & 0xffff;
    if (plat & PLAT_JUMPED) {
        if (res > 16.0 * res_prev) tau >>= 2;
        else if (res > 2.25 * res_prev) tau >>= 1;
        if (tau == 0) tau = PLAT_OFF;
    }
    const bool flat = fabs(res - res_prev) <= CVX_PLAT_TOL * res;
    int streak = flat ? ((plat >> 2) & 0xff) + 1 : 0;
    streak = streak > 255 ? 255 : streak;
    if (streak >= CVX_PLAT_N && tau != PLAT_OFF) {
        tau = tau == 0 ? CVX_PLAT_TAU0 : (2 * tau > CVX_PLAT_TAU_MAX ? CVX_PLAT_TAU_MAX : 2 * tau);
        plat = (tau << 10) | PLAT_JUMPED;
        return tau;
    }
    plat = (tau << 10) | (streak << 2);
    return 0;
}

// ---------------------------------------------------------------------------------
// Per-problem state machine.  The CUDA kernel is persistent: every lane pulls the
// next problem from a global counter as soon as its current one is finished, so
// no lane waits for the slowest problem of its warp / CTA (iteration counts vary
// from ~250 to 2500).  The three phases:
//     problem_begin   assembly, rho, Q/rho, start point
//     pass_dr         one DR iteration (returns whether the lane wants an Anderson step)
//     aa_step         warp-uniform Anderson extrapolation (history in tensor memory)
//     pass_eig        basis change + warm-started Jacobi sweep (or, once the DR loop
//                     has stopped, one more pass that polishes the
//                     eigen-decomposition); returns true when the problem is done
//     problem_finish  dual objective, optional Z, pose extraction
// V (100), M (55), T (55), L (10) are the problem's strided work arrays (220
// doubles, shared memory in the CUDA kernel); QR (45, Q/rho) is a strided scratch
// that may live in global memory.
// ---------------------------------------------------------------------------------
struct LaneState {
    double rho, dobj;
    double res_prev;   // squared residual of the previous DR iteration
    int32_t it;
    // phase 0: DR iterations; 1: polishing the eigen-decomposition of the last
    // (scaled) iterate; 2: eigen-decomposition of the unscaled Z (only when sigma != 1
    // and the solution may have rank > 1)
    int32_t phase;
    // While phase == 0 no code looks at `phase`, so its upper bits carry the plateau detector
    // (plateau_update; a register the 255-register solver kernel does not have to spare): every
    // `st.phase = ...` clears it.
    bool finite, iterating, converged;
    int32_t bad;   // tracked solver: consecutive DR iterations whose PSD projection was not certified (pnpl_track.cuh)
    AAState aa;
};

template <int S, class QRT>
CVX_HD void rescale_rho(Arr<S> V, Arr<S> M, Arr<S> L, QRT QR, LaneState& st, double factor)
{
    const double ic = 1.0 / factor;
#pragma unroll 1
    for (int j = 0; j < 10; ++j) {
        const double l = L[j];
        if (!(l < 0.0)) continue;
        const double dl = l * ic - l;
#pragma unroll 1
        for (int r = 0; r < 10; ++r) {
            const double a = dl * V[r * 10 + j];
#pragma unroll 1
            for (int c = 0; c <= r; ++c) M[sidx(r, c)] = fma(a, V[c * 10 + j], M[sidx(r, c)]);
        }
        L[j] = l * ic;
    }
#pragma unroll 1
    for (int e = 0; e < 45; ++e) QR[e] = QR[e] * ic;
    st.rho *= factor;
    aa_reset(st.aa);
    st.res_prev = 1e300;
}


// Assembly for one problem into the first 46 doubles of its record: Q/rho (45, packed) and rho.  Run by a
// lane-parallel pre-pass kernel (pre_kernel) so that the persistent solver's
// problem_begin -- which executes with a single active lane -- only has to copy.
// ... followed by the eigen-decomposition of the start point (V 100, lambda 10): 156 doubles.
constexpr int PRE_DOUBLES = 156;
constexpr int PRE_V = 46, PRE_L = 146;
#ifndef CVX_DUAL_GUESS
#define CVX_DUAL_GUESS 0.75
#endif
constexpr double DUAL_GUESS = CVX_DUAL_GUESS;
// ... and 0.9 with fewer than 8 points (lines only, minimal and near-minimal sets).  Host build with the tracked solver,
// 0.75 -> 0.9 (mean iterations; problems handed back): PnL-6 74.9 -> 69.5, 4.4 -> 1.4 %; 6 points 66.4 -> 62.4, 2.1 ->
// 0.6 %; 4 points 152 -> 139, 33 -> 22 %; 4 lines 226 -> 167 (12 -> 3 of 600 at the cap); 8+ points prefer 0.75
// (52.6 against 57.3).
CVX_HD double default_kappa(int n_pts)
{
#if defined(CVX_DUAL_GUESS_SMALL)
    return n_pts >= 8 ? DUAL_GUESS : CVX_DUAL_GUESS_SMALL;
#else
    return n_pts >= 8 ? DUAL_GUESS : 0.9;
#endif
}

template <class QOut>
CVX_HD void assemble_scaled(const Problem& pr, const Opts& o, QOut out)
{
    double Q[45], Bm[27];
    bool finite = assemble(pr.K, pr.pts_2d, pr.pts_3d, pr.n_pts, pr.line_2d, pr.line_3d, pr.n_lines, Q, Bm);
    double nq = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const double q = Q[sidx(i, j)];
            nq = fma((i == j) ? 1.0 : 2.0, q * q, nq);
        }
    const double rho = o.rho_rel * sqrt(nq);
    finite = finite && (rho > 0.0) && isfinite(rho);
    const double ir = 1.0 / rho;
#pragma unroll
    for (int e = 0; e < 45; ++e) out[e] = Q[e] * ir;
    out[45] = finite ? rho : nan("");
}

// Start point.  Primal: Z0 = blkdiag(I/3, sigma^2), feasible for the diagonal equalities.
// Dual: U0 = kappa Q/rho -- at the optimum rho U = Q - sum_k y_k P_k, so the cost matrix
// itself is the natural first guess (measured on seeded batches with kappa = 0.75: mean DR
// iterations 70.9 -> 60.5 on PnPL 8+4, 80.5 -> 62.2 on PnP-8, 169 -> 136 on PnL-6; kappa = 1
// is worse than none).  The DR state is M0 = Z0 - U0 = blkdiag(I/3 - kappa Q/rho, sigma^2);
// the solver needs its eigen-decomposition, which is that of Q: cold cyclic Jacobi, run
// lane-parallel in the pre-pass kernel (it would serialise inside the persistent kernel).
// `pre` holds Q/rho and rho on entry (assemble_scaled) and receives V, lambda.  V is a
// strided work array.
template <int S>
CVX_HD void start_decomposition(double* pre, const Opts& o, Arr<S> V)
{
    const bool finite = isfinite(pre[45]);
    double t[55];   // registers: every index below is a compile-time constant
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#pragma unroll
        for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double m = (i == j) ? (i == 9 ? o.sigma * o.sigma : 1.0 / 3.0) : 0.0;
            if (i < 9 && finite) m = fma(-o.kappa, pre[sidx(i, j)], m);
            t[sidx(i, j)] = m;
        }
    }
    if (finite && o.kappa != 0.0) {
        // the solver refines the decomposition by one warm-started sweep per iteration, so the
        // cold start only has to get close: stop after the sweep whose pivots were below 1e-2 of the
        // diagonal when it met them (it leaves them at ~1e-4).  Measured on the host build (3000 PnPL
        // 8+4 / 2000 PnP-8 / 2000 PnL-6 problems, tracked chain): mean DR iterations 52.37 / 53.58 / 69.98
        // with the former 1e-10 (two more sweeps), 52.36 / 53.66 / 69.82 with this one, 52.33 / 53.64 /
        // 69.96 with 1e-3; poses equal to 2e-9 rad.
#ifndef CVX_COLD_TOL
#define CVX_COLD_TOL 1e-4
#endif
#pragma unroll 1
        for (int s = 0; s < 10; ++s) {
            double dg = 0;
#pragma unroll
            for (int j = 0; j < 10; ++j) dg = fma(t[sidx(j, j)], t[sidx(j, j)], dg);
            if (!(jacobi_sweep_reg(t, V) > CVX_COLD_TOL * dg)) break;
        }
    }
#pragma unroll 4
    for (int e = 0; e < 100; ++e) pre[PRE_V + e] = V[e];
#pragma unroll
    for (int j = 0; j < 10; ++j) pre[PRE_L + j] = t[sidx(j, j)];
}

template <int S, class QRT>
CVX_HD void problem_begin(const double* pre, const Opts& o, Arr<S> V, Arr<S> M, Arr<S> L, QRT QR, LaneState& st)
{
    // Inside the persistent kernel this runs with ONE active lane while 31 wait, and the
    // record is cold (L2 / HBM): issue the loads in large independent groups so their
    // latencies overlap instead of adding up.
    const double rho = pre[45];
    const bool finite = isfinite(rho);
    {
        double q[45];
#pragma unroll
        for (int e = 0; e < 45; ++e) q[e] = pre[e];
        // M0 = blkdiag(I/3 - kappa Q/rho, sigma^2)
#pragma unroll
        for (int i = 0; i < 10; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const int e = sidx(i, j);
                double m = (i == j) ? (i == 9 ? o.sigma * o.sigma : 1.0 / 3.0) : 0.0;
                if (i < 9) {
                    QR[e] = q[e];
                    if (finite) m = fma(-o.kappa, q[e], m);
                }
                M[e] = m;
            }
    }
    // ... and its eigen-decomposition from the pre-pass
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double v[50];
#pragma unroll
        for (int e = 0; e < 50; ++e) v[e] = pre[PRE_V + 50 * h + e];
#pragma unroll
        for (int e = 0; e < 50; ++e) V[50 * h + e] = v[e];
    }
    {
        double l[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) l[j] = pre[PRE_L + j];
#pragma unroll
        for (int j = 0; j < 10; ++j) L[j] = l[j];
    }
    st.rho = rho;
    st.dobj = 0.0;
    st.phase = 0;
    st.it = 0;
    aa_reset(st.aa);
    st.res_prev = 1e300;
    st.finite = finite;
    st.iterating = finite;
    st.converged = false;
}

// ---------------------------------------------------------------------------------
// FP32 first phase (BASELINE.json configs[3]: "fp32 ADMM + fp64 extraction").  The
// DR iteration is self-correcting -- any M is a valid state -- so the iterations that
// only have to bring a problem from the cold start into the linear tail (||X - Z||_F
// from ~50 down to 0.15, about 25 of the ~55 iterations, no Anderson steps yet) can run
// in FP32: twice the FMA rate, half the shared memory per problem (so 256 instead of
// 128 problems per SM and two warps per scheduler instead of one).  The FP32 kernel
// exports (M, V, lambda, iteration count) per problem as doubles (WARM_DOUBLES); V is
// re-orthonormalised in FP64 (ortho_kernel; rotations preserve whatever V'V is, so an
// FP32-orthogonal basis would cap the accuracy at 1e-7 for good) and the FP64 solver
// continues from there to the full tolerance with Anderson acceleration.
// ---------------------------------------------------------------------------------
constexpr int WARM_DOUBLES = 166;   // M 55 | V 100 | lambda 10 | iterations

template <int S, class QRT>
CVX_HD void problem_begin32(const double* pre, const Opts& o, ArrT<S, float> V, ArrT<S, float> M, ArrT<S, float> L,
                            QRT QR)
{
    const bool finite = isfinite(pre[45]);
    const float s2 = (float)(o.sigma * o.sigma), kap = (float)o.kappa;
#pragma unroll 1
    for (int i = 0; i < 10; ++i)
#pragma unroll 1
        for (int j = 0; j <= i; ++j) {
            const int e = sidx(i, j);
            float m = (i == j) ? (i == 9 ? s2 : 1.f / 3.f) : 0.f;
            if (i < 9) {
                const float q = (float)pre[e];
                QR[e] = q;
                if (finite) m = fmaf(-kap, q, m);
            }
            M[e] = m;
        }
#pragma unroll 4
    for (int e = 0; e < 100; ++e) V[e] = (float)pre[PRE_V + e];
#pragma unroll 2
    for (int j = 0; j < 10; ++j) L[j] = (float)pre[PRE_L + j];
}

// one FP32 DR iteration + basis change + warm-started sweep; returns the squared residual
template <int S, class QRT>
CVX_HD float pass32(const Opts& o, ArrT<S, float> V, ArrT<S, float> M, ArrT<S, float> T, ArrT<S, float> L, QRT QR)
{
    float z[55];
    const float res = f32::dr_step(M, V, L, T, QR, (float)o.alpha, (float)(1.0 / o.sigma), (float)o.rowk, z);
    f32::rotate_into_basis(M, V, T);
    float t[55];
#pragma unroll
    for (int e = 0; e < 55; ++e) t[e] = T[e];
#pragma unroll 1
    for (int s = 0; s < o.sweeps; ++s) f32::jacobi_sweep_reg(t, V);
#pragma unroll
    for (int j = 0; j < 10; ++j) L[j] = t[sidx(j, j)];
    return res;
}

template <int S>
CVX_HD void problem_export32(ArrT<S, float> V, ArrT<S, float> M, ArrT<S, float> L, int it, double* w)
{
#pragma unroll 5
    for (int e = 0; e < 55; ++e) w[e] = (double)M[e];
#pragma unroll 4
    for (int e = 0; e < 100; ++e) w[55 + e] = (double)V[e];
#pragma unroll 2
    for (int e = 0; e < 10; ++e) w[155 + e] = (double)L[e];
    w[165] = (double)it;
}

// modified Gram-Schmidt on the columns of the exported eigenbasis, in FP64 (the basis
// is held in registers: every loop is unrolled, all indices are compile-time)
CVX_HD void warm_orthonormalise(double* w)
{
    double V[100];
#pragma unroll
    for (int e = 0; e < 100; ++e) V[e] = w[55 + e];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            if (k >= j) continue;
            double dt = 0;
#pragma unroll
            for (int i = 0; i < 10; ++i) dt = fma(V[i * 10 + j], V[i * 10 + k], dt);
#pragma unroll
            for (int i = 0; i < 10; ++i) V[i * 10 + j] = fma(-dt, V[i * 10 + k], V[i * 10 + j]);
        }
        double n = 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) n = fma(V[i * 10 + j], V[i * 10 + j], n);
        n = 1.0 / sqrt(n);
#pragma unroll
        for (int i = 0; i < 10; ++i) V[i * 10 + j] *= n;
    }
#pragma unroll
    for (int e = 0; e < 100; ++e) w[55 + e] = V[e];
}

// FP64 solver picks up a problem the FP32 phase has brought into the tail
template <int S, class QRT>
CVX_HD void problem_begin_warm(const double* pre, const double* w, const Opts& o, Arr<S> V, Arr<S> M, Arr<S> L, QRT QR,
                               LaneState& st)
{
    // loads in large independent groups (one active lane, cold records: see problem_begin)
    const double rho = pre[45];
    {
        double q[45];
#pragma unroll
        for (int e = 0; e < 45; ++e) q[e] = pre[e];
#pragma unroll
        for (int e = 0; e < 45; ++e) QR[e] = q[e];
    }
    {
        double m[55];
#pragma unroll
        for (int e = 0; e < 55; ++e) m[e] = w[e];
#pragma unroll
        for (int e = 0; e < 55; ++e) M[e] = m[e];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double v[55];
#pragma unroll
        for (int e = 0; e < 55; ++e) v[e] = w[55 + 55 * h + e];   // V (100) then lambda (10)
#pragma unroll
        for (int e = 0; e < 55; ++e) {
            const int k = 55 * h + e;
            if (k < 100) V[k] = v[e];
            else L[k - 100] = v[e];
        }
    }
    st.rho = rho;
    st.dobj = 0.0;
    st.phase = 0;
    st.it = (int32_t)w[165];
    aa_reset(st.aa);
    st.res_prev = 1e300;
    st.finite = isfinite(rho);
    st.iterating = st.finite && st.it < o.max_iters;
    if (st.finite && !st.iterating) st.phase = 1;
    st.converged = false;
}

// One loop body serves both the DR iterations (one warm-started sweep each) and the
// final passes that drive the eigen-decomposition of the last iterate to full
// convergence, so the sweep code exists once.  Returns true when done.
template <int S, class QR>
CVX_HD double dual_objective(Arr<S> V, Arr<S> L, QR qr, double rho, double sigma);

// Part 1 of a pass: one DR iteration (if the problem is still iterating).  Returns
// true when the lane wants an Anderson step on the iterate it just produced.
template <int S, class QRT>
CVX_HD bool pass_dr(const Opts& o, Arr<S> V, Arr<S> M, Arr<S> T, Arr<S> L, QRT QR, LaneState& st)
{
    if (!st.finite || !st.iterating) return false;
    double z[55];
    const double res = dr_step(M, V, L, T, QR, o.alpha, 1.0 / o.sigma, o.rowk, z);
    ++st.it;
#if defined(CVX_TRACE) && !defined(__CUDA_ARCH__)
    {
        int np_ = 0; double l1 = -1e300, l2 = -1e300, lneg = -1e300;
        for (int j = 0; j < 10; ++j) { const double l = L[j]; np_ += l > 0; if (l > l1) { l2 = l1; l1 = l; } else if (l > l2) l2 = l; if (l <= 0 && l > lneg) lneg = l; }
        printf("it %d res %.3e aa_mask %u npos %d l1 %.3e l2 %.3e lneg %.3e\n", st.it, sqrt(res), st.aa.mask, np_, l1, l2, lneg);
    }
#endif
    if (!(res > o.eps2)) {  // also leaves on NaN
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
    // plateau jump (see plateau_update): T holds the step g
    {
        const int tau = plateau_update(st.phase, res, st.res_prev);
        if (tau > 0) {
            const double ft = (double)tau;
#pragma unroll 1
            for (int e = 0; e < 55; ++e) M[e] = fma(ft, T[e], M[e]);
            aa_reset(st.aa);
            st.res_prev = res;
            return false;
        }
    }
    // Anderson acceleration only in the (locally linear) tail of the iteration:
    // extrapolating during the early active-set changes can throw M far away, from
    // where DR needs thousands of constant-length steps to walk back.  A residual
    // that grows by more than 2x after an accelerated step drops the history.
    const bool tail = res < o.aa_on2;
    if (!tail || res > 4.0 * st.res_prev) aa_reset(st.aa);
    st.res_prev = res;
    return tail;   // T holds the step g until the basis change in pass_eig
}

// Part 2 of a pass: basis change + warm-started Jacobi sweep; phase transitions.
// Returns true when the problem is done.
template <int S, class QRT>
CVX_HD bool pass_eig(const Opts& o, Arr<S> V, Arr<S> M, Arr<S> T, Arr<S> L, QRT QR, LaneState& st)
{
    if (!st.finite) return true;
    rotate_into_basis(M, V, T);
    double t[55];
#pragma unroll
    for (int e = 0; e < 55; ++e) t[e] = T[e];
    double off = 0.0, dg = 0.0;
#pragma unroll 1
    for (int s = 0; s < (st.iterating ? o.sweeps : 1); ++s) off = jacobi_sweep_reg(t, V);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const double l = t[sidx(j, j)];
        L[j] = l;
        dg = fma(l, l, dg);
    }
    if (st.iterating) {
        const double rf = rescale_factor(st.it);
        if (rf > 0.0) rescale_rho(V, M, L, QR, st, rf);
        return false;
    }
    if (!isfinite(dg)) return true;  // NaN-safe: a non-finite iterate ends the problem
    if (off > 1e-22 * dg) return false;
    if (st.phase == 2) return true;
    // phase 1 finished: V, L is the eigen-decomposition of the scaled iterate M'
    st.dobj = dual_objective(V, L, QR, st.rho, o.sigma);
    if (o.sigma == 1.0) return true;
    int above = 0;
#pragma unroll 1
    for (int j = 0; j < 10; ++j) above += (L[j] > 1e-3) ? 1 : 0;
    const double isig = 1.0 / o.sigma;
    if (above <= 1) {
        // rank <= 1 for certain (eigenvalues of Z lie in [lam'/sigma^2, lam']): the
        // eigenvector of Z is D^-1 v'; un-scaling row 9 of V is all extraction needs
#pragma unroll
        for (int j = 0; j < 10; ++j) V[90 + j] = V[90 + j] * isig;
        return true;
    }
    // possible rank > 1: the reference thresholds the eigenvalues of the UNSCALED Z
    // (cvxpnpl.py:499-502), so decompose Z = D^-1 P_psd(M') D^-1 itself; the passes
    // run in this same loop (M is free now)
#pragma unroll 1
    for (int r = 0; r < 10; ++r)
#pragma unroll 1
        for (int c = 0; c <= r; ++c) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < 10; ++j) s = fma(fmax(L[j], 0.0) * V[r * 10 + j], V[c * 10 + j], s);
            if (r == 9) s *= isig;
            if (c == 9) s *= isig;
            M[sidx(r, c)] = s;
        }
    st.phase = 2;
    return false;
}

// Dual objective.  At the fixed point Q + rho U = sum_k y_k P_k with U = V min(lam,0) V'
// the scaled dual slack, so for ANY affine-feasible point Zf the dual objective
// y_0 = sum_k y_k <P_k, Zf> = <Q + rho U, Zf>.  Zf = blkdiag(I/3, sigma^2) is used.
template <int S, class QR>
CVX_HD double dual_objective(Arr<S> V, Arr<S> L, QR qr, double rho, double sigma)
{
    double tr9 = 0, u99 = 0, tq = 0;
#pragma unroll 1
    for (int j = 0; j < 10; ++j) {
        const double ln = fmin(L[j], 0.0);
        double s = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) s = fma(V[i * 10 + j], V[i * 10 + j], s);
        tr9 = fma(ln, s, tr9);
        u99 = fma(ln, V[90 + j] * V[90 + j], u99);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) tq += qr[sidx(i, i)];
    // feasible point of the scaled problem: blkdiag(I/3, sigma^2)
    return rho * ((tq + tr9) * (1.0 / 3.0) + sigma * sigma * u99);
}

// A problem whose assembly is not finite is either non-finite input -- the reference carries the
// NaNs into SCS and returns the NaN pose (cvxpnpl.py:493-498): ST_NAN -- or finite input with an
// exactly singular 3x3 system (K, or the normal matrix N'N of degenerate bearings), where the
// reference's np.linalg.solve raises LinAlgError (cvxpnpl.py:37, 123-125, 548 | 579 | 623):
// ST_SINGULAR.  Only called for problems that are already known to be non-finite.
CVX_HD bool normal_system_singular(const Problem& pr)
{
    double Kl[9], Ki[9];
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        Kl[i] = pr.K[i];
        fin = fin && isfinite(Kl[i]);
    }
    inv3(Kl, Ki);
    bool ki_fin = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) ki_fin = ki_fin && isfinite(Ki[i]);
    double W[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < pr.n_pts; ++i) {
        double p[3];
        bearing(Ki, pr.pts_2d[2 * i], pr.pts_2d[2 * i + 1], p);
        fin = fin && isfinite(pr.pts_2d[2 * i]) && isfinite(pr.pts_2d[2 * i + 1]) && isfinite(pr.pts_3d[3 * i]) &&
              isfinite(pr.pts_3d[3 * i + 1]) && isfinite(pr.pts_3d[3 * i + 2]);
        const double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        W[0] += n2 - p[0] * p[0]; W[1] -= p[1] * p[0]; W[2] += n2 - p[1] * p[1];
        W[3] -= p[2] * p[0]; W[4] -= p[2] * p[1]; W[5] += n2 - p[2] * p[2];
    }
    for (int i = 0; i < pr.n_lines; ++i) {
        double a[3], b[3];
        bearing(Ki, pr.line_2d[4 * i], pr.line_2d[4 * i + 1], a);
        bearing(Ki, pr.line_2d[4 * i + 2], pr.line_2d[4 * i + 3], b);
        for (int e = 0; e < 4; ++e) fin = fin && isfinite(pr.line_2d[4 * i + e]);
        for (int e = 0; e < 6; ++e) fin = fin && isfinite(pr.line_3d[6 * i + e]);
        double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        if (!(nn > 0.0)) continue;   // coincident endpoints: NaN normal in the reference, not an exception
        const double inv = 2.0 / nn;  // both endpoints of the line add n n'
        W[0] += inv * n[0] * n[0]; W[1] += inv * n[1] * n[0]; W[2] += inv * n[1] * n[1];
        W[3] += inv * n[2] * n[0]; W[4] += inv * n[2] * n[1]; W[5] += inv * n[2] * n[2];
    }
    if (!fin) return false;          // non-finite data: NaN pose, like the reference
    if (!ki_fin) return true;        // singular K
    const double det = W[0] * (W[2] * W[5] - W[4] * W[4]) - W[1] * (W[1] * W[5] - W[4] * W[3]) +
                       W[3] * (W[1] * W[4] - W[2] * W[3]);
    return isfinite(det) && !isfinite(1.0 / det);
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

// result of a problem whose 3x3 system is exactly singular: no pose, ST_SINGULAR (LinAlgError in the scalar API)
CVX_HD void singular_result(double* R_out, double* t_out, Result& rs)
{
    const double qnan = nan("");
#pragma unroll 4
    for (int i = 0; i < 36; ++i) R_out[i] = qnan;
#pragma unroll 4
    for (int i = 0; i < 12; ++i) t_out[i] = qnan;
    rs.n_poses = 0;
    rs.status = ST_SINGULAR;
    rs.pobj = qnan;
    rs.dobj = qnan;
}

template <int S, class QRT>
CVX_HD void problem_finish(const Problem& pr, const Opts& o, Arr<S> V, Arr<S> M, Arr<S> T, Arr<S> L, QRT QR,
                           const LaneState& st, double* R_out, double* t_out, double* Z_out, Result& rs)
{
    double lam[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) lam[j] = L[j];
    int32_t status = ST_NAN;
    if (st.finite) {
        status = st.converged ? ST_OK : ST_MAX_ITERS;
#pragma unroll
        for (int j = 0; j < 10; ++j)
            if (!isfinite(lam[j])) status = ST_NAN;
    }
    const double dobj = (status != ST_NAN) ? st.dobj : nan("");
    if (Z_out) write_Z(V, lam, status == ST_NAN, Z_out);
    if (status == ST_NAN && normal_system_singular(pr)) {
        singular_result(R_out, t_out, rs);
        rs.iters = st.it;
        return;
    }
    // Q and B are re-assembled (cheap) into the now free M / T regions
    Arr<S> Qs = M;
    Arr<S> Bs = T;
    if (status != ST_NAN) assemble(pr.K, pr.pts_2d, pr.pts_3d, pr.n_pts, pr.line_2d, pr.line_3d, pr.n_lines, Qs, Bs);
    double pobj;
    rs.n_poses = extract_poses(V, lam, Qs, Bs, status, dobj, sqrt(o.eps2), R_out, t_out, pobj);
    rs.status = status;
    rs.iters = st.it;
    rs.pobj = pobj;
    rs.dobj = dobj;
}

// Deferred finish.  Inside the persistent kernel a finishing lane runs alone (the
// other 31 lanes of its warp wait), so everything expensive at the end of a problem
// -- re-assembly, the 3x3 SVD projections, the multi-solution recovery -- is NOT
// done there: the lane only parks its eigen-decomposition (V 100, lambda 10, dual
// objective, preliminary status = 112 doubles) in global memory, and a second, fully
// lane-parallel kernel (extract_parked) turns it into poses.  ncu: the in-loop
// finish cost 13 % of the warp-issue samples before the split.
constexpr int PARK_DOUBLES = 112;

template <int S>
CVX_HD void problem_park(Arr<S> V, Arr<S> L, const LaneState& st, double* park, int32_t* iters_out)
{
    int32_t status = ST_NAN;
    if (st.finite) {
        status = st.converged ? ST_OK : ST_MAX_ITERS;
#pragma unroll 1
        for (int j = 0; j < 10; ++j)
            if (!isfinite(L[j])) status = ST_NAN;
    }
#pragma unroll 4
    for (int e = 0; e < 100; ++e) park[e] = V[e];
#pragma unroll 2
    for (int j = 0; j < 10; ++j) park[100 + j] = L[j];
    park[110] = (status != ST_NAN) ? st.dobj : nan("");
    park[111] = (double)status;
    *iters_out = st.it;
}

// second half of the finish: from the parked state to poses.  V is a strided work
// array (100) that receives the parked eigenvectors, Qs (45) / Bs (27) receive the
// re-assembled problem.
template <int SV, int S>
CVX_HD void extract_parked(const Problem& pr, const Opts& o, const double* park, Arr<SV> V, Arr<S> Qs, Arr<S> Bs,
                           double* R_out, double* t_out, double* Z_out, Result& rs)
{
    double lam[10];
    if (V.p != park) {
#pragma unroll 4
        for (int e = 0; e < 100; ++e) V[e] = park[e];
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) lam[j] = park[100 + j];
    const double dobj = park[110];
    int32_t status = (int32_t)park[111];
    if (Z_out) write_Z(V, lam, status == ST_NAN, Z_out);
    if (status == ST_NAN && normal_system_singular(pr)) {
        singular_result(R_out, t_out, rs);
        return;
    }
    if (status != ST_NAN) assemble(pr.K, pr.pts_2d, pr.pts_3d, pr.n_pts, pr.line_2d, pr.line_3d, pr.n_lines, Qs, Bs);
    double pobj;
    rs.n_poses = extract_poses(V, lam, Qs, Bs, status, dobj, sqrt(o.eps2), R_out, t_out, pobj);
    rs.status = status;
    rs.pobj = pobj;
    rs.dobj = dobj;
}

// The whole path for one problem, sequentially (host harness).
template <int S, class QRT, class Hist>
CVX_HD void solve_problem(const Problem& pr, const Opts& o, Arr<S> V, Arr<S> M, Arr<S> T, Arr<S> L, QRT QR,
                          const Hist& H, double* R_out, double* t_out, double* Z_out, Result& rs)
{
    LaneState st;
    double pre[PRE_DOUBLES];
    assemble_scaled(pr, o, pre);
    start_decomposition(pre, o, V);
    problem_begin(pre, o, V, M, L, QR, st);
    int wslot = 0;
#pragma unroll 1
    for (int guard = 0; guard < o.max_iters + 40; ++guard) {
        const bool want = pass_dr(o, V, M, T, L, QR, st);
        if (H.any(want)) aa_step(M, T, H, st.aa, want, wslot, (float)st.res_prev);
        wslot = (wslot + 1 == AA_M) ? 0 : wslot + 1;
        if (pass_eig(o, V, M, T, L, QR, st)) break;
    }
    problem_finish(pr, o, V, M, T, L, QR, st, R_out, t_out, Z_out, rs);
}

}  // namespace cvx
