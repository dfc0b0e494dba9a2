// pnpl_warp.cuh -- warp-per-problem Douglas-Rachford iterations: the low-latency
// path for the stragglers of a batch.
//
// Why it exists.  The thread-per-problem solver (pnpl_solve.cuh) is throughput
// optimal, but one pass of one problem is a 14 k-instruction dependent stream that
// takes ~26 us whatever else the SM does.  Iteration counts have a long tail (PnPL
// 8+4: median 69, p99 108, max ~800 of 1e5; PnP / PnL batches contain problems that
// run into the 2500-iteration cap), so once the work queue is empty the whole GPU
// waits for a handful of lanes: measured, a 1e5 PnPL batch took 19.4 ms of which the
// balanced bulk is ~11 ms, and a PnP-8 batch took 65 ms = 2500 x 26 us.
//
// What it does.  When the queue of the persistent kernel has run dry, every lane
// still iterating after a grace period parks its complete DR state (M, V, lambda,
// Q/rho, rho, iteration count: one HAND slab entry) and leaves.  This kernel then
// gives every such problem a whole WARP: the 10x10 matrices sit in shared memory in
// full (unpacked) form, the 32 lanes split every step by matrix entry (Z = V L+ V',
// the affine projection: one lane per equality, T = V'MV), the Jacobi sweep of a DR
// iteration takes all 45 rotation angles from T as it is and applies them to
// register-resident rows of V and M V (warp_sweep_stale; the cold decomposition of a
// handed-back problem uses the exact round-by-round sweep, warp_sweep: the five disjoint
// rotations of a round at once as 25 independent two-sided 2x2 block updates plus 50 row
// updates of V), and the Anderson dot products are spread over 28 lanes.  One iteration
// costs ~7 k cycles (3.55 us, measured on an idle GPU) instead of ~26 us.  The
// algorithm, its constants and its stopping rule are those of the thread path; only
// the Anderson history (FP32, unscaled, in shared memory) restarts at the hand-over.
// When the DR loop of a problem has stopped the state goes back into its slab entry
// and the thread kernel (resume mode) polishes the eigen-decomposition and parks it
// for finish_kernel exactly as it does for its own problems.
#pragma once

#include "pnpl_core.cuh"
#include "pnpl_solve.cuh"

namespace cvx {

// ---- hand-over slab entry (doubles) ---------------------------------------------
constexpr int HO_M = 0;        // 55 packed DR iterate
constexpr int HO_V = 55;       // 100 eigenbasis
constexpr int HO_L = 155;      // 10 eigenvalues
constexpr int HO_Q = 165;      // 45 packed Q / rho
constexpr int HO_RHO = 210;
constexpr int HO_IT = 211;
constexpr int HO_FLAGS = 212;  // bit 0: converged; bit 1: DR loop finished (by a warp-per-problem kernel)
constexpr int HO_B = 213;      // problem index
constexpr int HAND_DOUBLES = 216;

template <int S, class QRT>
CVX_HD void problem_handoff(Arr<S> V, Arr<S> M, Arr<S> L, QRT QR, const LaneState& st, int64_t b, double* h)
{
#pragma unroll 5
    for (int e = 0; e < 55; ++e) h[HO_M + e] = M[e];
#pragma unroll 4
    for (int e = 0; e < 100; ++e) h[HO_V + e] = V[e];
#pragma unroll 2
    for (int e = 0; e < 10; ++e) h[HO_L + e] = L[e];
#pragma unroll 5
    for (int e = 0; e < 45; ++e) h[HO_Q + e] = QR[e];
    h[HO_RHO] = st.rho;
    h[HO_IT] = (double)st.it;
    h[HO_FLAGS] = 0.0;
    h[HO_B] = (double)b;
}

// back into the thread solver: DR loop finished, eigen-decomposition to be polished
template <int S, class QRT>
CVX_HD int64_t problem_resume(const double* h, Arr<S> V, Arr<S> M, Arr<S> L, QRT QR, LaneState& st)
{
#pragma unroll 5
    for (int e = 0; e < 55; ++e) M[e] = h[HO_M + e];
#pragma unroll 4
    for (int e = 0; e < 100; ++e) V[e] = h[HO_V + e];
#pragma unroll 2
    for (int e = 0; e < 10; ++e) L[e] = h[HO_L + e];
#pragma unroll 5
    for (int e = 0; e < 45; ++e) QR[e] = h[HO_Q + e];
    st.rho = h[HO_RHO];
    st.it = (int32_t)h[HO_IT];
    st.converged = ((int)h[HO_FLAGS] & 1) != 0;   // (bit 1: the DR loop was finished by a warp)
    st.dobj = 0.0;
    st.phase = 1;
    st.finite = true;
    st.iterating = false;
    st.res_prev = 1e300;
    aa_reset(st.aa);
    return (int64_t)h[HO_B];
}

#if defined(__CUDACC__)

#define CVX_TRI_ROW(i0, j0, s0, i1, j1, s1, i2, j2, s2, ROW) {i0, j0, s0, i1, j1, s1, i2, j2, s2, ROW},
__constant__ signed char c_triples[15][10] = {CVX_TRIPLES(CVX_TRI_ROW)};
#undef CVX_TRI_ROW

struct alignas(16) WarpSmem {
    double M[100], V[100], T[100], X[100], Z[100], Q[100];   // full 10x10, row-major
    double L[10], cs[10];
    double csall[90];   // (c, s) of the 45 pivot pairs of a sweep in round order (warp_sweep_stale)
    float gp[56], sp[56], gk[56];
    float dG[AA_M][56], dS[AA_M][56];
    float gram[AA_GRAM_WORDS], dots[16];
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float warp_sumf(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// (row, column) of packed lower-triangular index p
__device__ __forceinline__ void unpack_idx(int p, int& r, int& c)
{
    r = 0;
#pragma unroll
    for (int k = 1; k < 10; ++k) r += (p >= (k * (k + 1)) / 2) ? 1 : 0;
    c = p - (r * (r + 1)) / 2;
}

// pivot pair k (0..4) of round r (0..8): circle method, player 9 fixed
__device__ __forceinline__ void round_pair(int r, int k, int& p, int& q)
{
    int a, b;
    if (k == 0) {
        a = r;
        b = 9;
    } else {
        a = r + k;
        a = a >= 9 ? a - 9 : a;
        b = r + 9 - k;
        b = b >= 9 ? b - 9 : b;
    }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

// Pivot pairs of a lane's work items for all nine rounds of a sweep, four bits each:
// T block (pa, qa) x (pb, qb); V items rows lane / 5 (pair pb, qb) and (lane + 32) / 5 (pair p1, q1)
__device__ __forceinline__ void sweep_tables(int lane, uint32_t pk[9])
{
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        int pa, qa, pb, qb, p1, q1;
        round_pair(r, lane < 25 ? lane / 5 : 0, pa, qa);
        round_pair(r, lane % 5, pb, qb);
        round_pair(r, (lane + 32) % 5, p1, q1);
        pk[r] = (uint32_t)pa | ((uint32_t)qa << 4) | ((uint32_t)pb << 8) | ((uint32_t)qb << 12) | ((uint32_t)p1 << 16) |
                ((uint32_t)q1 << 20);
    }
}

// One Jacobi sweep on S.T (full form) accumulating the rotations into S.V: 9 rounds of 5 disjoint rotations.
// Round: lanes 0..4 compute (c, s); then 25 lanes update one 2x2 block of T each (T' = J'TJ, blocks are
// disjoint -> in place) and all lanes rotate rows of V.
__device__ __forceinline__ void warp_sweep(WarpSmem& S, int lane, const uint32_t pk[9])
{
    const int ka = lane < 25 ? lane / 5 : 0, kb = lane % 5, k1 = (lane + 32) % 5;
    const int i0 = lane / 5, i1 = (lane + 32) / 5;
#pragma unroll 1   // rolled: measured faster than nine unrolled copies (pk then lives in local memory)
    for (int round = 0; round < 9; ++round) {
        const int pa = pk[round] & 15, qa = (pk[round] >> 4) & 15, pb = (pk[round] >> 8) & 15,
                  qb = (pk[round] >> 12) & 15, p1 = (pk[round] >> 16) & 15, q1 = (pk[round] >> 20) & 15;
        if (lane < 5) {
            // lane k < 5 owns pair k = (pb, qb)
            double c, s;
            jacobi_cs_fast(S.T[pb * 11], S.T[qb * 11], S.T[qb * 10 + pb], c, s);
            S.cs[2 * lane] = c;
            S.cs[2 * lane + 1] = s;
        }
        __syncwarp();
        const double cb = S.cs[2 * kb], sb = S.cs[2 * kb + 1];
        if (lane < 25) {
            const double ca = S.cs[2 * ka], sa = S.cs[2 * ka + 1];
            const double b00 = S.T[pa * 10 + pb], b01 = S.T[pa * 10 + qb], b10 = S.T[qa * 10 + pb],
                         b11 = S.T[qa * 10 + qb];
            const double y00 = ca * b00 - sa * b10, y01 = ca * b01 - sa * b11;
            const double y10 = sa * b00 + ca * b10, y11 = sa * b01 + ca * b11;
            const bool dg = ka == kb;   // the pivot itself is annihilated exactly
            S.T[pa * 10 + pb] = y00 * cb - y01 * sb;
            S.T[pa * 10 + qb] = dg ? 0.0 : y00 * sb + y01 * cb;
            S.T[qa * 10 + pb] = dg ? 0.0 : y10 * cb - y11 * sb;
            S.T[qa * 10 + qb] = y10 * sb + y11 * cb;
        }
        {
            const double vp = S.V[i0 * 10 + pb], vq = S.V[i0 * 10 + qb];
            S.V[i0 * 10 + pb] = fma(cb, vp, -sb * vq);
            S.V[i0 * 10 + qb] = fma(sb, vp, cb * vq);
        }
        if (lane < 18) {
            const double c = S.cs[2 * k1], s = S.cs[2 * k1 + 1];
            const double vp = S.V[i1 * 10 + p1], vq = S.V[i1 * 10 + q1];
            S.V[i1 * 10 + p1] = fma(c, vp, -s * vq);
            S.V[i1 * 10 + q1] = fma(s, vp, c * vq);
        }
        __syncwarp();
    }
}

// The sweep of a DR iteration (warm start: T = V'MV is nearly diagonal) with all 45 rotation angles taken from T AS IT IS
// when the sweep begins, instead of from the partly rotated matrix round by round.  The rotations are still exact
// (orthogonal to double precision) and applied in the cyclic order, so V stays orthonormal; only the annihilation of
// the pivots is first-order instead of exact: the off-diagonal part left behind is O(off^2 / gap), the same order a
// cyclic sweep leaves, and the next iteration recomputes T from scratch anyway.  Measured on the host build (the same
// change in jacobi_sweep_reg, 300..3000 problems per family, tools/stale_angles_host.py): mean DR iterations PnPL 8+4
// 52.8 -> 52.6, PnL-6 67.5 -> 68.1, PnP-8 53.8 -> 53.8, 4 points 121.4 -> 121.3, same poses (1e-9 rad), same problems at
// the cap.  What it buys: the nine DEPENDENT rounds of (angle chain ~270 cycles + 2x2 block update + two warp barriers,
// ~690 cycles each: half of the warp's iteration, profiles/r2bf) become one round of angles for all pairs on 32 + 13
// lanes, the 45 rotations applied to register-resident rows of V and of M V by twenty lanes, and the eigenvalues as
// column dot products of the two -- T itself is not rotated at all.
// On entry S.T = V' M V (full form) and S.X = M V; on exit S.V is rotated (S.X with it) and S.L holds the new eigenvalue
// estimates.
__device__ __forceinline__ void warp_sweep_stale(WarpSmem& S, int lane)
{
    // ---- angles: pivot pair number i = 5 * round + k, lanes take i = lane and i = lane + 32 ----------------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < 45) {
            int p, q;
            round_pair(i / 5, i % 5, p, q);
            double c, s;
            jacobi_cs_fast(S.T[p * 11], S.T[q * 11], S.T[q * 10 + p], c, s);
            S.csall[2 * i] = c;
            S.csall[2 * i + 1] = s;
        }
    }
    __syncwarp();
    // ---- the 45 rotations in cyclic order on the rows of V (lanes 0..9) and, alike, on the rows of X = M V (lanes
    //      10..19; step 5 left it there): a row lives in registers, every pivot pair is a compile-time constant, so the
    //      nine rounds need no barrier and no address arithmetic; X V-rotated is M V_new, which the eigenvalues need.
    if (lane < 20) {
        double* row = lane < 10 ? &S.V[lane * 10] : &S.X[(lane - 10) * 10];
        double v[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) v[j] = row[j];
        const double2* cs2 = reinterpret_cast<const double2*>(S.csall);
#pragma unroll
        for (int round = 0; round < 9; ++round)
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                int p, q;
                round_pair(round, k, p, q);
                const double2 cs = cs2[5 * round + k];
                const double vp = v[p], vq = v[q];
                v[p] = fma(cs.x, vp, -cs.y * vq);
                v[q] = fma(cs.y, vp, cs.x * vq);
            }
#pragma unroll
        for (int j = 0; j < 10; ++j) row[j] = v[j];
    }
    __syncwarp();
    // ---- eigenvalue estimates: lambda_j = v_j' M v_j = column j of V_new dot column j of M V_new -----------------------
    if (lane < 10) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 10; k += 2) {
            s0 = fma(S.V[k * 10 + lane], S.X[k * 10 + lane], s0);
            s1 = fma(S.V[(k + 1) * 10 + lane], S.X[(k + 1) * 10 + lane], s1);
        }
        S.L[lane] = s0 + s1;
    }
    __syncwarp();
}

// Cold eigen-decomposition of S.M (full form) by one warp: V = I, T = M, cyclic sweeps.  The rotations annihilate a
// pivot to ~1e-7 of its size (jacobi_cs_fast), so the off-diagonal part shrinks at least that fast per sweep once the
// quadratic phase is over; ten sweeps leave it far below what the solver's own per-iteration sweep / the polishing
// passes of the thread solver expect as a warm start.
__device__ __forceinline__ void warp_cold_decompose(WarpSmem& S, int lane, const uint32_t pk[9])
{
    for (int e = lane; e < 100; e += 32) {
        S.T[e] = S.M[e];
        S.V[e] = (e / 10 == e % 10) ? 1.0 : 0.0;
    }
    __syncwarp();
#pragma unroll 1
    for (int sw = 0; sw < 10; ++sw) warp_sweep(S, lane, pk);
    if (lane < 10) S.L[lane] = S.T[lane * 11];
    __syncwarp();
}

// Normal equations of the Anderson step: packed Gram matrix (lower), right-hand side,
// valid-column mask -> coefficients.  Same arithmetic as aa_step.
__device__ __forceinline__ bool aa_solve_packed(const float* gram, const float* rg, uint32_t mask, float f[AA_M])
{
#define CVX_TI(i, j) (((i) * ((i) + 1)) / 2 + (j))
    float A[AA_GRAM_WORDS], r[AA_M];
    float tr = 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) {
        const bool vi = (mask >> i) & 1u;
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
            if (j <= i) A[CVX_TI(i, j)] = (vi && ((mask >> j) & 1u)) ? gram[CVX_TI(i, j)] : 0.f;
        r[i] = vi ? rg[i] : 0.f;
        tr += A[CVX_TI(i, i)];
    }
    bool pd = tr > 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) A[CVX_TI(i, i)] += 1e-6f * tr + (((mask >> i) & 1u) ? 0.f : 1.f);
#pragma unroll
    for (int j = 0; j < AA_M; ++j) {
        float d = A[CVX_TI(j, j)];
#pragma unroll
        for (int k = 0; k < AA_M; ++k)
            if (k < j) d = fmaf(-A[CVX_TI(j, k)], A[CVX_TI(j, k)], d);
        pd = pd && (d > 0.f);
        const float id = rsqrtf(pd ? d : 1.f);
        A[CVX_TI(j, j)] = id;
#pragma unroll
        for (int i = 0; i < AA_M; ++i) {
            if (i <= j) continue;
            float t = A[CVX_TI(i, j)];
#pragma unroll
            for (int k = 0; k < AA_M; ++k)
                if (k < j) t = fmaf(-A[CVX_TI(i, k)], A[CVX_TI(j, k)], t);
            A[CVX_TI(i, j)] = t * id;
        }
    }
#pragma unroll
    for (int i = 0; i < AA_M; ++i) {
        float t = r[i];
#pragma unroll
        for (int k = 0; k < AA_M; ++k)
            if (k < i) t = fmaf(-A[CVX_TI(i, k)], r[k], t);
        r[i] = t * A[CVX_TI(i, i)];
    }
#pragma unroll
    for (int ii = 0; ii < AA_M; ++ii) {
        const int i = AA_M - 1 - ii;
        float t = r[i];
#pragma unroll
        for (int k = 0; k < AA_M; ++k)
            if (k > i) t = fmaf(-A[CVX_TI(k, i)], r[k], t);
        r[i] = t * A[CVX_TI(i, i)];
    }
#undef CVX_TI
    bool ok = mask != 0u && pd;
#pragma unroll
    for (int j = 0; j < AA_M; ++j) ok = ok && isfinite(r[j]);
#pragma unroll
    for (int j = 0; j < AA_M; ++j) f[j] = (ok && ((mask >> j) & 1u)) ? r[j] : 0.f;
    return ok;
}

// Equality group g (0..14) of the affine projection as one lane applies it: the three entries it touches, their
// coefficients and 1 / |a|^2 (times rowk for the row groups of the rc variant).
struct AffLane {
    int e0, e1, e2;
    double s0, s1, a2, k;
};
__device__ __forceinline__ void aff_lane_init(AffLane& a, int g, const Opts& o, double isig, double inrm9)
{
    const signed char* t = c_triples[g];
    a.e0 = t[0] * 10 + t[1];
    a.e1 = t[3] * 10 + t[4];
    a.e2 = t[6] * 10 + t[7];
    a.s0 = t[2];
    a.s1 = t[5];
    a.a2 = (t[6] == 9) ? t[8] * isig : (double)t[8];
    a.k = (t[6] == 9) ? inrm9 : (t[9] ? o.rowk * (1.0 / 3.0) : (1.0 / 3.0));
}

// DR iterations of one problem by one warp.  S holds M, V, L, Q (= Q/rho, full form
// with a zero last row/column); `it` continues the problem's iteration count; the loop ends on
// convergence or at iteration it_stop (<= o.max_iters).
__device__ __noinline__ void warp_dr_loop(WarpSmem& S, const Opts& o, int lane, int& it, bool& converged,
                                          double& rho, int it_stop)
{
    const unsigned FULL = 0xffffffffu;
    const double isig = 1.0 / o.sigma, inrm9 = 1.0 / (2.0 + isig * isig);
    // packed entries owned by this lane: p = lane and p = lane + 32 (< 55)
    int er[2], ec[2];
    unpack_idx(lane, er[0], ec[0]);
    unpack_idx(lane + 32 < 55 ? lane + 32 : 0, er[1], ec[1]);
    const int np = lane + 32 < 55 ? 2 : 1;
    uint32_t pk[9];
    sweep_tables(lane, pk);
    uint32_t mask = 0u;
    bool have_prev = false;
    int wslot = 0;
    double res_prev = 1e300;
    int32_t plat = 0;   // plateau detector (warp-uniform; restarts at the hand-over)
    converged = false;
    // equality group of this lane (step 2), read from constant memory ONCE: inside the loop the fifteen lanes' table
    // reads are fifteen serialised constant-cache accesses per entry (ncu, profiles/r2bf: ~1 k cycles per iteration)
    AffLane af;
    aff_lane_init(af, lane < 15 ? lane : 0, o, isig, inrm9);
    for (;;) {
        // ---- 1. Z = V max(L,0) V',  W = 2 Z - M - Q/rho  (W into X) -------------------
        {
            double lp[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) lp[j] = fmax(S.L[j], 0.0);
            _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
                const int r = er[q], c = ec[q];
                double z0 = 0.0, z1 = 0.0;
#pragma unroll
                for (int j = 0; j < 10; j += 2) {
                    // only the positive eigenpairs contribute (two or three of ten as a rule): warp-uniform branches,
                    // and skipping a term that is exactly zero leaves the sum as it was
                    if (lp[j] > 0.0) z0 = fma(lp[j] * S.V[r * 10 + j], S.V[c * 10 + j], z0);
                    if (lp[j + 1] > 0.0) z1 = fma(lp[j + 1] * S.V[r * 10 + j + 1], S.V[c * 10 + j + 1], z1);
                }
                const double z = z0 + z1;
                const double w = 2.0 * z - S.M[r * 10 + c] - S.Q[r * 10 + c];
                S.Z[r * 10 + c] = z;
                S.X[r * 10 + c] = w;
                S.X[c * 10 + r] = w;
            }
        }
        __syncwarp();
        // ---- 2. X = P_aff(W): one lane per equality group (15 triples, 9 diagonal entries,
        //         the homogeneous corner).  Inputs first, then a barrier, then the stores.
        double x0 = 0, x1 = 0, x2 = 0;
        int e0 = 0, e1 = 0, e2 = 0;
        if (lane < 15) {
            e0 = af.e0;
            e1 = af.e1;
            e2 = af.e2;
            const double w0 = S.X[e0], w1 = S.X[e1], w2 = S.X[e2];
            const double rr = (af.s0 * w0 + af.s1 * w1 + af.a2 * w2) * af.k;
            x0 = w0 - af.s0 * rr;
            x1 = w1 - af.s1 * rr;
            x2 = w2 - af.a2 * rr;
        } else if (lane >= 16 && lane < 25) {
            const int i = lane - 16, cc = i / 3, rr = i % 3;
            double R = 0, C = 0, G = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double wk = S.X[k * 11];
                G += wk;
                if (k % 3 == rr) R += wk;
                if (k / 3 == cc) C += wk;
            }
            x0 = S.X[i * 11] - o.rowk * (R - 1.0) * (1.0 / 3.0) - (C - 1.0) * (1.0 / 3.0) + o.rowk * (G - 3.0) * (1.0 / 9.0);
            e0 = i * 11;
        }
        __syncwarp();
        if (lane < 15) {
            S.X[e0] = x0;
            S.X[e1] = x1;
            S.X[e2] = x2;
        } else if (lane >= 16 && lane < 25) {
            S.X[e0] = x0;
        } else if (lane == 25) {
            S.X[99] = o.sigma * o.sigma;
        }
        __syncwarp();
        // ---- 3. M += alpha (X - Z); the step g = alpha (X - Z) replaces Z; residual -----
        double rs = 0.0;
        _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
            const int r = er[q], c = ec[q];
            const double d = S.X[r * 10 + c] - S.Z[r * 10 + c];
            const double m = fma(o.alpha, d, S.M[r * 10 + c]);
            S.M[r * 10 + c] = m;
            S.M[c * 10 + r] = m;
            S.Z[r * 10 + c] = o.alpha * d;
            rs = fma((r == c) ? 1.0 : 2.0, d * d, rs);
        }
        const double res = warp_sum(rs);
        ++it;
        if (!(res > o.eps2)) {   // also leaves on NaN
            converged = (res <= o.eps2);
            break;
        }
        if (it >= it_stop) break;   // o.max_iters, or earlier when the caller only lends the warp for a while
        __syncwarp();
        // ---- 4. Anderson acceleration in the tail (same rules as pass_dr / aa_step) ------
        // plateau jump (same rule as pass_dr: plateau_update); Z holds the step g
        const int tau = o.anderson ? plateau_update(plat, res, res_prev) : 0;
        if (tau > 0) {
            const double ft = (double)tau;
            _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
                const int r = er[q], c = ec[q];
                const double m = fma(ft, S.Z[r * 10 + c], S.M[r * 10 + c]);
                S.M[r * 10 + c] = m;
                S.M[c * 10 + r] = m;
            }
            mask = 0u;
            have_prev = false;
            res_prev = res;
            __syncwarp();
        } else if (o.anderson) {
            const bool tail = res < o.aa_on2;
            if (!tail || res > 4.0 * res_prev) {
                mask = 0u;
                have_prev = false;
            }
            res_prev = res;
            if (tail) {
                const bool close = have_prev;
                _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
                    const int p = lane + 32 * q;
                    const float gf = (float)S.Z[er[q] * 10 + ec[q]];
                    const float d = gf - S.gp[p];
                    S.dG[wslot][p] = close ? d : 0.f;
                    S.dS[wslot][p] = close ? S.sp[p] + d : 0.f;
                    S.gk[p] = gf;
                }
                mask = close ? (mask | (1u << wslot)) : (mask & ~(1u << wslot));
                __syncwarp();
                // 14 dot products (col_j . g, col_j . col_wslot), each split over two lanes
                {
                    const int k = lane >> 1, h = lane & 1;
                    float acc = 0.f;
                    if (k < 2 * AA_M) {
                        const float* a = S.dG[k < AA_M ? k : k - AA_M];
                        const float* bv = (k < AA_M) ? S.gk : S.dG[wslot];
                        const int lo = h * 28, hi = h ? 55 : 28;
                        for (int p = lo; p < hi; ++p) acc = fmaf(a[p], bv[p], acc);
                    }
                    acc += __shfl_xor_sync(FULL, acc, 1);
                    if (k < 2 * AA_M && h == 0) S.dots[k] = acc;
                }
                __syncwarp();
                if (lane < AA_M) {
                    const int i = lane > wslot ? lane : wslot, j = lane > wslot ? wslot : lane;
                    S.gram[(i * (i + 1)) / 2 + j] = S.dots[AA_M + lane];
                }
                __syncwarp();
                float gr[AA_GRAM_WORDS], rg[AA_M], f[AA_M];
#pragma unroll
                for (int e = 0; e < AA_GRAM_WORDS; ++e) gr[e] = S.gram[e];
#pragma unroll
                for (int j = 0; j < AA_M; ++j) rg[j] = S.dots[j];
                const bool ok = aa_solve_packed(gr, rg, mask, f);
                float adj[2] = {0.f, 0.f}, ng = 0.f, ns = 0.f;
                _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
                    const int p = lane + 32 * q;
#pragma unroll
                    for (int j = 0; j < AA_M; ++j) adj[q] = fmaf(f[j], S.dS[j][p], adj[q]);
                    const float gf = S.gk[p], stp = gf - adj[q];
                    ng = fmaf(gf, gf, ng);
                    ns = fmaf(stp, stp, ns);
                }
                ng = warp_sumf(ng);
                ns = warp_sumf(ns);
                const bool apply = ok && (ns <= AA_MAX_STEP2 * ng);
                _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
                    const int p = lane + 32 * q, r = er[q], c = ec[q];
                    const float gf = S.gk[p];
                    if (apply) {
                        const double m = S.M[r * 10 + c] - (double)adj[q];
                        S.M[r * 10 + c] = m;
                        S.M[c * 10 + r] = m;
                    }
                    S.gp[p] = gf;
                    S.sp[p] = apply ? gf - adj[q] : gf;
                }
                if (mask != 0u && !apply) mask = 0u;
                have_prev = true;
                __syncwarp();
            }
        }
        wslot = (wslot + 1 == AA_M) ? 0 : wslot + 1;
        // ---- 5. T = V' M V  (M V into X, then the lower triangle of V' (M V), mirrored) ---
        for (int e = lane; e < 100; e += 32) {
            const int r = e / 10, c = e - 10 * r;
            double s0 = 0.0, s1 = 0.0;   // (two chains: half the dependent latency)
#pragma unroll
            for (int k = 0; k < 10; k += 2) {
                s0 = fma(S.M[r * 10 + k], S.V[k * 10 + c], s0);
                s1 = fma(S.M[r * 10 + k + 1], S.V[(k + 1) * 10 + c], s1);
            }
            S.X[e] = s0 + s1;
        }
        __syncwarp();
        _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
            const int r = er[q], c = ec[q];
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < 10; k += 2) {
                s0 = fma(S.V[k * 10 + r], S.X[k * 10 + c], s0);
                s1 = fma(S.V[(k + 1) * 10 + r], S.X[(k + 1) * 10 + c], s1);
            }
            const double s = s0 + s1;
            S.T[r * 10 + c] = s;
            S.T[c * 10 + r] = s;
        }
        __syncwarp();
        // ---- 6. one Jacobi sweep: angles from T as it is (warp_sweep_stale); -DCVX_WARP_EXACT_SWEEP: round by round -------
#ifdef CVX_WARP_EXACT_SWEEP
        warp_sweep(S, lane, pk);
        if (lane < 10) S.L[lane] = S.T[lane * 11];
        __syncwarp();
#else
        warp_sweep_stale(S, lane);
#endif
        // ---- 7. slow problem: continue with a smaller penalty, once (rescale_rho) ---------
        const double rf = rescale_factor(it);
        if (rf > 0.0) {
            const double ic = 1.0 / rf;
            _Pragma("unroll") for (int q = 0; q < 2; ++q) if (q < np) {
                const int r = er[q], c = ec[q];
                double m = S.M[r * 10 + c];
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const double l = S.L[j];
                    m = fma((l < 0.0 ? l * ic - l : 0.0) * S.V[r * 10 + j], S.V[c * 10 + j], m);
                }
                S.M[r * 10 + c] = m;
                S.M[c * 10 + r] = m;
            }
            for (int e = lane; e < 100; e += 32) S.Q[e] *= ic;
            __syncwarp();
            if (lane < 10 && S.L[lane] < 0.0) S.L[lane] *= ic;
            rho *= rf;
            mask = 0u;
            have_prev = false;
            res_prev = 1e300;
            __syncwarp();
        }
    }
    __syncwarp();
}

#endif  // __CUDACC__

}  // namespace cvx
