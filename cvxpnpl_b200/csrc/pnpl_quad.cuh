// pnpl_quad.cuh -- FOUR warps per problem: the lowest-latency form of the Douglas-Rachford
// iteration, for the last problems of a batch.
//
// Why it exists.  The warp-per-problem kernel (pnpl_warp.cuh) takes ~6.2 us per iteration
// however idle the GPU is: one warp issues all ~4400 instructions of an iteration itself
// (ncu, profiles/r2bb: 41 % of its stall samples are fixed-latency dependencies, 22 % the
// shared-memory / MUFU scoreboard, the warp issues every ~4.6 cycles).  Once only a few
// problems are left -- the two or three of 1e5 PnL-6 problems that run into the
// 2500-iteration cap, the slowest few hundred of a degenerate sweep -- the step lasts as
// long as THEIR remaining iterations times that latency (PnL-6: 13.5 of 23 ms).  With
// nothing else to do the SMs can spend more threads per problem:
//   * every phase of the iteration that is parallel over matrix entries (Z = V L+ V',
//     W, the M update, M V and V'(M V), the Anderson history and its fourteen dot
//     products) is spread over 128 threads, one entry / one eighth of a dot product each;
//   * the Jacobi sweep is split by ROLE: warp 0 walks the nine rounds on T alone
//     (rotation angles on five lanes, the 25 two-sided 2x2 block updates), publishing
//     every round's (c, s) in shared memory; warps 1 and 2 apply the rounds to the rows
//     of V (rows 0-5 and 6-9, five lanes per row) as they are published, i.e. the 50 row
//     updates per round leave the critical path;
//   * the small serial pieces (7x7 normal equations of the Anderson step) are computed
//     redundantly by all threads, which saves a broadcast.
// Same algorithm, constants, stopping rule and shared-memory layout (WarpSmem) as the warp
// kernel; like there, the Anderson history restarts when a problem arrives.
#pragma once

#include "pnpl_warp.cuh"

namespace cvx {

#if defined(__CUDACC__)

constexpr int QUAD_NT = 128;

struct QuadSmem {
    WarpSmem W;
    double cs[9][10];    // (c, s) of the five pivot pairs of each round of the current sweep
    double red[4];       // cross-warp sums
    float redf[4][2];
    int flag;            // rounds of the current sweep whose rotations are published
    int pad;
};

// One sweep: warp 0 on T, warps 1-2 on V.  Ends with a CTA barrier.
__device__ __forceinline__ void quad_sweep(QuadSmem& Q, int tid, const uint32_t pk[9])
{
    WarpSmem& S = Q.W;
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        const int ka = lane < 25 ? lane / 5 : 0, kb = lane % 5;
#pragma unroll 1
        for (int round = 0; round < 9; ++round) {
            const int pa = pk[round] & 15, qa = (pk[round] >> 4) & 15, pb = (pk[round] >> 8) & 15,
                      qb = (pk[round] >> 12) & 15;
            volatile double* cs = Q.cs[round];
            // this lane's 2x2 block: loaded before the angles are known (the loads overlap that chain)
            const double b00 = S.T[pa * 10 + pb], b01 = S.T[pa * 10 + qb], b10 = S.T[qa * 10 + pb], b11 = S.T[qa * 10 + qb];
            if (lane < 5) {
                double c, s;
                jacobi_cs_fast(S.T[pb * 11], S.T[qb * 11], S.T[qb * 10 + pb], c, s);
                cs[2 * lane] = c;
                cs[2 * lane + 1] = s;
            }
            __syncwarp();
            if (lane == 31) {   // (a lane without a block of T: the fence stays off the critical path)
                __threadfence_block();
                *(volatile int*)&Q.flag = round + 1;   // warps 1-2 may apply this round to V
            }
            if (lane < 25) {
                const double ca = cs[2 * ka], sa = cs[2 * ka + 1], cb = cs[2 * kb], sb = cs[2 * kb + 1];
                const double y00 = ca * b00 - sa * b10, y01 = ca * b01 - sa * b11;
                const double y10 = sa * b00 + ca * b10, y11 = sa * b01 + ca * b11;
                const bool dg = ka == kb;   // the pivot itself is annihilated exactly
                S.T[pa * 10 + pb] = y00 * cb - y01 * sb;
                S.T[pa * 10 + qb] = dg ? 0.0 : y00 * sb + y01 * cb;
                S.T[qa * 10 + pb] = dg ? 0.0 : y10 * cb - y11 * sb;
                S.T[qa * 10 + qb] = y10 * sb + y11 * cb;
            }
            __syncwarp();
        }
        if (lane < 10) S.L[lane] = S.T[lane * 11];
    } else if (warp <= 2) {
        // rows of V: warp 1 rows 0..5 (30 lanes), warp 2 rows 6..9 (20 lanes); the five lanes of a row take the five
        // disjoint pivot pairs of a round, rounds are separated by a warp barrier
        const int row = (warp == 1 ? 0 : 6) + lane / 5, k = lane % 5;
        const bool act = lane < (warp == 1 ? 30 : 20);
#pragma unroll 1
        for (int round = 0; round < 9; ++round) {
            const int p = (pk[round] >> 8) & 15, q = (pk[round] >> 12) & 15;   // pair lane % 5 (sweep_tables)
            if (lane == 0) {
                while (*(volatile int*)&Q.flag <= round) {
                }
                __threadfence_block();
            }
            __syncwarp();
            if (act) {
                const volatile double* cs = Q.cs[round];
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double vp = S.V[row * 10 + p], vq = S.V[row * 10 + q];
                S.V[row * 10 + p] = fma(c, vp, -s * vq);
                S.V[row * 10 + q] = fma(s, vp, c * vq);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid == 0) Q.flag = 0;   // (the next sweep is several CTA barriers away)
}

// DR iterations of one problem by four warps.  Q.W holds M, V, L, Q (= Q/rho, full form with a zero last
// row/column) and a cleared Anderson history; `it` continues the problem's iteration count; the loop ends on convergence
// or at iteration it_stop (<= o.max_iters).  Every thread of the CTA calls this with the same arguments and returns
// the same it / converged / rho.
__device__ __noinline__ void quad_dr_loop(QuadSmem& Q, const Opts& o, int tid, int& it, bool& converged, double& rho,
                                          int it_stop)
{
    WarpSmem& S = Q.W;
    const unsigned FULL = 0xffffffffu;
    const int warp = tid >> 5, lane = tid & 31;
    const double isig = 1.0 / o.sigma, inrm9 = 1.0 / (2.0 + isig * isig);
    const bool own = tid < 55;   // packed entry p = tid of the symmetric matrices
    int er, ec;
    unpack_idx(own ? tid : 0, er, ec);
    uint32_t pk[9];
    sweep_tables(lane, pk);
    // warp-uniform (indeed CTA-uniform) solver state, kept redundantly by every thread
    uint32_t mask = 0u;
    bool have_prev = false;
    int wslot = 0;
    double res_prev = 1e300;
    int32_t plat = 0;
    converged = false;
    AffLane af;   // this thread's equality group (step 2), out of the loop like in warp_dr_loop
    aff_lane_init(af, tid < 15 ? tid : 0, o, isig, inrm9);
    if (tid == 0) Q.flag = 0;
    __syncthreads();
    for (;;) {
        // ---- 1. Z = V max(L,0) V',  W = 2 Z - M - Q/rho  (W into X) -------------------
        if (own) {
            double z0 = 0.0, z1 = 0.0;
#pragma unroll
            for (int j = 0; j < 10; j += 2) {
                const double l0 = fmax(S.L[j], 0.0), l1 = fmax(S.L[j + 1], 0.0);
                if (l0 > 0.0) z0 = fma(l0 * S.V[er * 10 + j], S.V[ec * 10 + j], z0);
                if (l1 > 0.0) z1 = fma(l1 * S.V[er * 10 + j + 1], S.V[ec * 10 + j + 1], z1);
            }
            const double z = z0 + z1;
            const double w = 2.0 * z - S.M[er * 10 + ec] - S.Q[er * 10 + ec];
            S.Z[er * 10 + ec] = z;
            S.X[er * 10 + ec] = w;
            S.X[ec * 10 + er] = w;
        }
        __syncthreads();
        // ---- 2. X = P_aff(W): one thread per equality group; inputs, barrier, stores ------------
        double x0 = 0, x1 = 0, x2 = 0;
        int e0 = 0, e1 = 0, e2 = 0;
        if (tid < 15) {
            e0 = af.e0;
            e1 = af.e1;
            e2 = af.e2;
            const double w0 = S.X[e0], w1 = S.X[e1], w2 = S.X[e2];
            const double rr = (af.s0 * w0 + af.s1 * w1 + af.a2 * w2) * af.k;
            x0 = w0 - af.s0 * rr;
            x1 = w1 - af.s1 * rr;
            x2 = w2 - af.a2 * rr;
        } else if (tid >= 32 && tid < 41) {   // (second warp: runs beside the triples)
            const int i = tid - 32, cc = i / 3, rr = i % 3;
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
        __syncthreads();
        if (tid < 15) {
            S.X[e0] = x0;
            S.X[e1] = x1;
            S.X[e2] = x2;
        } else if (tid >= 32 && tid < 41) {
            S.X[e0] = x0;
        } else if (tid == 64) {
            S.X[99] = o.sigma * o.sigma;
        }
        __syncthreads();
        // ---- 3. M += alpha (X - Z); the step g = alpha (X - Z) replaces Z; residual -----
        double rs = 0.0;
        if (own) {
            const double d = S.X[er * 10 + ec] - S.Z[er * 10 + ec];
            const double m = fma(o.alpha, d, S.M[er * 10 + ec]);
            S.M[er * 10 + ec] = m;
            S.M[ec * 10 + er] = m;
            S.Z[er * 10 + ec] = o.alpha * d;
            rs = ((er == ec) ? 1.0 : 2.0) * d * d;
        }
        if (warp < 2) {
            rs = warp_sum(rs);
            if (lane == 0) Q.red[warp] = rs;
        }
        __syncthreads();
        const double res = Q.red[0] + Q.red[1];
        ++it;
        if (!(res > o.eps2)) {   // also leaves on NaN
            converged = (res <= o.eps2);
            break;
        }
        if (it >= it_stop) break;
        // ---- 4. plateau jump / Anderson acceleration (same rules as warp_dr_loop) ------
        const int tau = o.anderson ? plateau_update(plat, res, res_prev) : 0;
        if (tau > 0) {
            if (own) {
                const double m = fma((double)tau, S.Z[er * 10 + ec], S.M[er * 10 + ec]);
                S.M[er * 10 + ec] = m;
                S.M[ec * 10 + er] = m;
            }
            mask = 0u;
            have_prev = false;
            res_prev = res;
            __syncthreads();
        } else if (o.anderson) {
            const bool tail = res < o.aa_on2;
            if (!tail || res > 4.0 * res_prev) {
                mask = 0u;
                have_prev = false;
            }
            res_prev = res;
            if (tail) {
                const bool close = have_prev;
                if (own) {
                    const float gf = (float)S.Z[er * 10 + ec];
                    const float d = gf - S.gp[tid];
                    S.dG[wslot][tid] = close ? d : 0.f;
                    S.dS[wslot][tid] = close ? S.sp[tid] + d : 0.f;
                    S.gk[tid] = gf;
                }
                mask = close ? (mask | (1u << wslot)) : (mask & ~(1u << wslot));
                __syncthreads();
                // 14 dot products (col_j . g, col_j . col_wslot), eight threads each
                {
                    const int k = tid >> 3, h = tid & 7;
                    float acc = 0.f;
                    if (k < 2 * AA_M) {
                        const float* a = S.dG[k < AA_M ? k : k - AA_M];
                        const float* bv = (k < AA_M) ? S.gk : S.dG[wslot];
#pragma unroll
                        for (int i = 0; i < 7; ++i) {
                            const int p = h + 8 * i;
                            if (p < 55) acc = fmaf(a[p], bv[p], acc);
                        }
                    }
                    acc += __shfl_xor_sync(FULL, acc, 1);
                    acc += __shfl_xor_sync(FULL, acc, 2);
                    acc += __shfl_xor_sync(FULL, acc, 4);
                    if (k < 2 * AA_M && h == 0) S.dots[k] = acc;
                }
                __syncthreads();
                if (tid < AA_M) {
                    const int i = tid > wslot ? tid : wslot, j = tid > wslot ? wslot : tid;
                    S.gram[(i * (i + 1)) / 2 + j] = S.dots[AA_M + tid];
                }
                __syncthreads();
                float gr[AA_GRAM_WORDS], rg[AA_M], f[AA_M];
#pragma unroll
                for (int e = 0; e < AA_GRAM_WORDS; ++e) gr[e] = S.gram[e];
#pragma unroll
                for (int j = 0; j < AA_M; ++j) rg[j] = S.dots[j];
                const bool ok = aa_solve_packed(gr, rg, mask, f);
                float adj = 0.f, ng = 0.f, ns = 0.f;
                if (own) {
#pragma unroll
                    for (int j = 0; j < AA_M; ++j) adj = fmaf(f[j], S.dS[j][tid], adj);
                    const float gf = S.gk[tid], stp = gf - adj;
                    ng = gf * gf;
                    ns = stp * stp;
                }
                if (warp < 2) {
                    ng = warp_sumf(ng);
                    ns = warp_sumf(ns);
                    if (lane == 0) {
                        Q.redf[warp][0] = ng;
                        Q.redf[warp][1] = ns;
                    }
                }
                __syncthreads();
                ng = Q.redf[0][0] + Q.redf[1][0];
                ns = Q.redf[0][1] + Q.redf[1][1];
                const bool apply = ok && (ns <= AA_MAX_STEP2 * ng);
                if (own) {
                    const float gf = S.gk[tid];
                    if (apply) {
                        const double m = S.M[er * 10 + ec] - (double)adj;
                        S.M[er * 10 + ec] = m;
                        S.M[ec * 10 + er] = m;
                    }
                    S.gp[tid] = gf;
                    S.sp[tid] = apply ? gf - adj : gf;
                }
                if (mask != 0u && !apply) mask = 0u;
                have_prev = true;
                __syncthreads();
            }
        }
        wslot = (wslot + 1 == AA_M) ? 0 : wslot + 1;
        // ---- 5. T = V' M V  (M V into X, then the lower triangle of V' (M V), mirrored) ---
        if (tid < 100) {
            const int r = tid / 10, c = tid - 10 * r;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < 10; k += 2) {
                s0 = fma(S.M[r * 10 + k], S.V[k * 10 + c], s0);
                s1 = fma(S.M[r * 10 + k + 1], S.V[(k + 1) * 10 + c], s1);
            }
            S.X[tid] = s0 + s1;
        }
        __syncthreads();
        if (own) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < 10; k += 2) {
                s0 = fma(S.V[k * 10 + er], S.X[k * 10 + ec], s0);
                s1 = fma(S.V[(k + 1) * 10 + er], S.X[(k + 1) * 10 + ec], s1);
            }
            const double s = s0 + s1;
            S.T[er * 10 + ec] = s;
            S.T[ec * 10 + er] = s;
        }
        __syncthreads();
        // ---- 6. one Jacobi sweep -------------------------------------------------------
        quad_sweep(Q, tid, pk);
        // ---- 7. slow problem: continue with a smaller penalty, once (rescale_rho) ---------
        const double rf = rescale_factor(it);
        if (rf > 0.0) {
            const double ic = 1.0 / rf;
            if (own) {
                double m = S.M[er * 10 + ec];
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const double l = S.L[j];
                    m = fma((l < 0.0 ? l * ic - l : 0.0) * S.V[er * 10 + j], S.V[ec * 10 + j], m);
                }
                S.M[er * 10 + ec] = m;
                S.M[ec * 10 + er] = m;
            }
            if (tid < 100) S.Q[tid] *= ic;
            __syncthreads();
            if (tid < 10 && S.L[tid] < 0.0) S.L[tid] *= ic;
            rho *= rf;
            mask = 0u;
            have_prev = false;
            res_prev = 1e300;
            __syncthreads();
        }
    }
    __syncthreads();
}

#endif  // __CUDACC__

}  // namespace cvx
