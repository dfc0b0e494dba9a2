// pnpl_track2.cuh -- the tracked solver of pnpl_track.cuh with TWO threads per problem.
//
// Why.  The thread-per-problem solver holds 128 problems per SM -- that is what shared memory (178 doubles each) and
// tensor memory (the Anderson history, 512 words each) have room for -- so it runs ONE warp per scheduler, and with
// one warp a dependent FP64 instruction issues every 8.2 cycles (measured, tools/micro/dfma_latency.cu; rsqrt 75):
// ncu shows 4.4 cycles per issued instruction, 23 % of the issue slots used.  More problems do not fit; more threads
// per problem do.  Here every problem is owned by a PAIR of threads in different warps (warp w and warp w + 4 of a
// 256-thread CTA; both see the same tensor-memory lanes) that split every step of an iteration:
//     DR step          thread A: the first eight triples; thread B: the other seven and the diagonal equalities
//     Anderson step    A: entries 8c .. 8c+3 of every 8-entry chunk c; B: entries 8c+4 .. 8c+7 (each keeps its half of
//                      the history in its own 256 tensor-memory columns); A solves the 7x7 normal equations
//     eigenpair step   both reduce M with the two reflectors (redundantly: no exchange); A factorises (s0 I - B) and
//                      solves for the first pair, B does (s1 I - B) [and the certificate at shift 0] and the second;
//                      both do the 2x2 Rayleigh-Ritz; each assembles its own vector
// with the state in shared memory and a named barrier of the two warps between the phases (five to eight per
// iteration).  Two warps per scheduler, each walking about half of the instruction stream.
//
// The phases are ordinary __host__ __device__ functions parameterised by the thread's ROLE, so the host build
// (tests/host) runs them role after role between the points where the kernel has its barriers.
#pragma once

#include "pnpl_track.cuh"

namespace cvx {

// ---- exchange scratch of a problem: strided doubles, then strided floats ----------------------------------------
constexpr int X2_RES = 0;   // 2: partial squared residuals of A and B
constexpr int X2_CTL = 2;   // 1: A's decision before a pass: >= 0 a new problem, -1 nothing, -3 give the current one up
constexpr int X2_OK = 3;    // 2: certificate flags (A: slot 0; B: slot 1 and, if both are positive, shift 0)
constexpr int X2_DOUBLES = 5;
constexpr int XF_PART = 0;    // 9 floats: B's partial dot products nd_0..6, ndd, ndg
constexpr int XF_COEF = 9;    // 9 words: seven coefficient words (FP16 pairs), 1 / fs, ok
constexpr int XF_NORM = 18;   // 4 floats: |g|^2, |step|^2 partials of A, of B
constexpr int X2_FLOATS = 22;
// what the two halves of the eigenpair step hand each other, in the (then free) step array G:
constexpr int XG_X = 0;       // 16: x_0 (A), x_1 (B)
constexpr int XG_V1 = 16;     // 10: reflector 1
constexpr int XG_V2 = 26;     // 9: reflector 2 (its tenth component is zero)
constexpr int XG_CT = 35;     // 16: c~_0, c~_1
constexpr int XG_TH = 51;     // 4: th~_0, th~_1, jc, js        (G[55] stays the Anderson step's zero pad)

// ---- Anderson history of one role: 256 tensor-memory columns -----------------------------------------------------
constexpr int A2_GP = 0;      // 7 chunks x 4: g_{k-1} (FP32)
constexpr int A2_SP = 28;     // 7 x 2: previous step (FP16 pairs, scaled)
constexpr int A2_DG = 42;     // 7 x [7 columns x 2]: dG
constexpr int A2_DS = 140;    // 7 x [7 columns x 2]: dM + dG
constexpr int A2_END = 238;
constexpr int A2_ROLE_WORDS = 256;
constexpr int A2_GRAM_LO = A2_END;                   // 18 words: Gram entries 0..17 (role A's spare columns)
constexpr int A2_GRAM_HI = A2_ROLE_WORDS + A2_END;   // 18 words: Gram entries 18..27, rg_0..6, pad (role B's spare columns)

// the 14 words (7 columns x 2) of chunk c of a role's column block
template <class Hist>
CVX_HD void aa2_ld_cols(const Hist& H, int off, uint32_t w[14])
{
    H.template ld<8>(off, w);
    H.template ld<4>(off + 8, w + 8);
    H.template ld<2>(off + 12, w + 12);
}

// Part A of the Anderson step for one role: close the newest column on the role's entries, partial dot products.
template <int S, class Hist>
CVX_HD void aa2_part_a(int role, Arr<S> G, const Hist& H, const AAState& aa, bool active, int wslot, float res2,
                       float& sc_out, float nd[AA_M], float& ndd_out, float& ndg_out)
{
    const int base = role * A2_ROLE_WORDS;
    const bool close = active && aa.have_prev;
    const float sc = close ? fminf(aa_scale(res2), 4.f * aa.scale_prev) : aa_scale(res2);
    const float ratio = sc / aa.scale_prev;
    float ndd = 0.f, ndg = 0.f;
#pragma unroll
    for (int j = 0; j < AA_M; ++j) nd[j] = 0.f;
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t gpw[4], spw[2], cw[14], dgw[2], dsw[2];
        H.template ld<4>(base + A2_GP + 4 * c, gpw);
        H.template ld<2>(base + A2_SP + 2 * c, spw);
        aa2_ld_cols(H, base + A2_DG + 14 * c, cw);
        H.wait_ld();
        float gf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) gf[u] = (float)G[c * 8 + 4 * role + u];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            float s0, s1, r0, r1;
            unpack_h2(spw[w], s0, s1);
            const float d0 = (gf[2 * w] - w2f(gpw[2 * w])) * sc, d1 = (gf[2 * w + 1] - w2f(gpw[2 * w + 1])) * sc;
            dgw[w] = close ? pack_h2(d0, d1) : 0u;
            dsw[w] = close ? pack_h2(fmaf(s0, ratio, d0), fmaf(s1, ratio, d1)) : 0u;
            unpack_h2(dgw[w], r0, r1);
            ndd = fmaf(r0, r0, fmaf(r1, r1, ndd));
            ndg = fmaf(r0, gf[2 * w], fmaf(r1, gf[2 * w + 1], ndg));
        }
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 2; ++w) nd[j] = hfma_hh(cw[2 * j + w], dgw[w], hfma_ll(cw[2 * j + w], dgw[w], nd[j]));
        H.template st<2>(base + A2_DG + 14 * c + 2 * wslot, dgw);
        H.template st<2>(base + A2_DS + 14 * c + 2 * wslot, dsw);
    }
    sc_out = sc;
    ndd_out = ndd;
    ndg_out = ndg;
}

// Role A only: Gram matrix / right-hand side update and the 7x7 normal equations (the arithmetic of aa_step).
// nd, ndd, ndg are the TOTALS; `mask` is the column mask after this step's update.  Writes the coefficient words,
// 1 / fs and the ok flag to cf[0..8].
template <class Hist>
CVX_HD void aa2_solve(const Hist& H, uint32_t mask, bool active, int wslot, float sc, const float nd[AA_M], float ndd,
                      float ndg, uint32_t cf[9])
{
    float gram[AA_GRAM_WORDS], rg[AA_M];
    {
        uint32_t lo[18], hi[18];
        H.template ld<16>(A2_GRAM_LO, lo);
        H.template ld<2>(A2_GRAM_LO + 16, lo + 16);
        H.template ld<16>(A2_GRAM_HI, hi);
        H.template ld<2>(A2_GRAM_HI + 16, hi + 16);
        H.wait_ld();
#pragma unroll
        for (int i = 0; i < AA_M; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const int e = (i * (i + 1)) / 2 + j;
                float v = w2f(e < 18 ? lo[e] : hi[e - 18]);
                if (i == wslot && j == wslot) v = ndd;
                else if (i == wslot) v = nd[j];
                else if (j == wslot) v = nd[i];
                gram[e] = v;
                if (e < 18) lo[e] = f2w(v);
                else hi[e - 18] = f2w(v);
            }
        const float isc = 1.f / sc;
#pragma unroll
        for (int j = 0; j < AA_M; ++j) {
            const float prev = w2f(hi[10 + j]);
            rg[j] = (j == wslot) ? ndg : fmaf(nd[j], isc, prev);
            hi[10 + j] = f2w(rg[j]);
        }
        hi[17] = 0u;
        H.template st<16>(A2_GRAM_LO, lo);
        H.template st<2>(A2_GRAM_LO + 16, lo + 16);
        H.template st<16>(A2_GRAM_HI, hi);
        H.template st<2>(A2_GRAM_HI + 16, hi + 16);
    }
    float A[AA_GRAM_WORDS], r[AA_M];
    float tr = 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) {
        const bool vi = (mask >> i) & 1u;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const bool vj = (mask >> j) & 1u;
            A[(i * (i + 1)) / 2 + j] = (vi && vj) ? gram[(i * (i + 1)) / 2 + j] : 0.f;
        }
        r[i] = vi ? rg[i] : 0.f;
        tr += A[(i * (i + 1)) / 2 + i];
    }
    bool pd = tr > 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) A[(i * (i + 1)) / 2 + i] += 1e-6f * tr + (((mask >> i) & 1u) ? 0.f : 1.f);
#define CVX_TI(i, j) (((i) * ((i) + 1)) / 2 + (j))
#pragma unroll
    for (int j = 0; j < AA_M; ++j) {
        float d = A[CVX_TI(j, j)];
#pragma unroll
        for (int k = 0; k < AA_M; ++k)
            if (k < j) d = fmaf(-A[CVX_TI(j, k)], A[CVX_TI(j, k)], d);
        pd = pd && (d > 0.f);
        const float id = f32::cvx_rsqrt(pd ? d : 1.f);
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
    bool ok = active && mask != 0u && pd;
    float fmx = 0.f;
#pragma unroll
    for (int j = 0; j < AA_M; ++j) {
        ok = ok && isfinite(r[j]);
        fmx = fmaxf(fmx, fabsf(r[j]));
    }
    int e = (int)((f2w(fmx) >> 23) & 0xffu) - 126;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    const float fs = w2f((uint32_t)(127 - e) << 23);
#pragma unroll
    for (int j = 0; j < AA_M; ++j) {
        const float f = (ok && ((mask >> j) & 1u)) ? r[j] * fs : 0.f;
        cf[j] = pack_h2(f, f);
    }
    cf[7] = (uint32_t)(127 + e) << 23;   // 1 / fs
    cf[8] = ok ? 1u : 0u;
    H.wait_st();
}

// Part B for one role: extrapolate on the role's entries, remember g_k and the step, partial norms.
template <int S, class Hist>
CVX_HD void aa2_part_b(int role, Arr<S> M, Arr<S> G, const Hist& H, bool active, const uint32_t cf[9], float sc,
                       float& ng_out, float& ns_out)
{
    const int base = role * A2_ROLE_WORDS;
    const float ifs = w2f(cf[7]);
    const bool ok = cf[8] != 0u;
    float ng = 0.f, ns = 0.f;
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t cw[14], gkw[4], spw[2];
        aa2_ld_cols(H, base + A2_DS + 14 * c, cw);
        H.wait_ld();
        float adj[4], stp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) adj[u] = 0.f;
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                adj[2 * w] = hfma_ll(cw[2 * j + w], cf[j], adj[2 * w]);
                adj[2 * w + 1] = hfma_hl(cw[2 * j + w], cf[j], adj[2 * w + 1]);
            }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = c * 8 + 4 * role + u;
            adj[u] *= ifs;
            const float gf = active ? (float)G[e] : 0.f;
            gkw[u] = f2w(gf);
            stp[u] = gf - adj[u];
            if (ok && e < 55) M[e] -= (double)adj[u];
            ng = fmaf(gf, gf, ng);
            ns = fmaf(stp[u], stp[u], ns);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) spw[w] = pack_h2(stp[2 * w] * sc, stp[2 * w + 1] * sc);
        H.template st<4>(base + A2_GP + 4 * c, gkw);
        H.template st<2>(base + A2_SP + 2 * c, spw);
    }
    H.wait_st();
    ng_out = ng;
    ns_out = ns;
}

// Part C (rare) for one role: the extrapolated step was rejected -- undo it on the role's entries, keep the plain step.
template <int S, class Hist>
CVX_HD void aa2_part_c(int role, Arr<S> M, const Hist& H, bool reject, const uint32_t cf[9], float sc)
{
    const int base = role * A2_ROLE_WORDS;
    const float ifs = w2f(cf[7]);
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t cw[14], gkw[4], spw[2];
        aa2_ld_cols(H, base + A2_DS + 14 * c, cw);
        H.template ld<4>(base + A2_GP + 4 * c, gkw);
        H.template ld<2>(base + A2_SP + 2 * c, spw);
        H.wait_ld();
        float adj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) adj[u] = 0.f;
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                adj[2 * w] = hfma_ll(cw[2 * j + w], cf[j], adj[2 * w]);
                adj[2 * w + 1] = hfma_hl(cw[2 * j + w], cf[j], adj[2 * w + 1]);
            }
        if (reject) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = c * 8 + 4 * role + u;
                if (e < 55) M[e] += (double)(adj[u] * ifs);
            }
#pragma unroll
            for (int w = 0; w < 2; ++w) spw[w] = pack_h2(w2f(gkw[2 * w]) * sc, w2f(gkw[2 * w + 1]) * sc);
        }
        H.template st<2>(base + A2_SP + 2 * c, spw);
    }
    H.wait_st();
}

// Cholesky of  shift I - B  with B (8x8, packed lower) in REGISTERS, and the solve (shift I - B) x = rhs.
CVX_HD bool chol8_solve_reg(const double b[36], double shift, const double rhs[8], double x[8])
{
    double l[36];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) l[sidx(i, j)] = ((i == j) ? shift : 0.0) - b[sidx(i, j)];
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
    return ok;
}

// First half of the eigenpair step (track_step, steps 1-3) for one role.  Both roles reduce M with the two reflectors;
// role A leaves the reflectors, c~ and the 2x2 Ritz data in G for the second half; each role factorises and solves its
// own slot.  `run` = this thread owns a live problem (the others just pass through).
template <int S>
CVX_HD void t2_step_p1(int role, Arr<S> M, Arr<S> U, Arr<S> G, Arr<S> X)
{
    double v1[10], v2[10], beta1, beta2;
    bool bad_reflector = false;
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
        beta1 = 1.0 / (nr1 * (nr1 + a9));
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
        bad_reflector = !(n1 > 0.0) || !(n2 > 0.0) || !isfinite(beta1) || !isfinite(beta2);
    }
    double m[55];
#pragma unroll
    for (int e = 0; e < 55; ++e) m[e] = M[e];
    reflect_sym<10>(m, v1, beta1);
    reflect_sym<9>(m, v2, beta2);
    double ct0[8], ct1[8], th0, th1, jc, js, jt;
    jacobi_cs(m[sidx(8, 8)], m[sidx(9, 9)], m[sidx(9, 8)], jc, js, jt);
    th0 = fma(-jt, m[sidx(9, 8)], m[sidx(8, 8)]);
    th1 = fma(jt, m[sidx(9, 8)], m[sidx(9, 9)]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double c8 = m[sidx(8, i)], c9 = m[sidx(9, i)];
        ct0[i] = fma(jc, c8, -js * c9);
        ct1[i] = fma(js, c8, jc * c9);
    }
    if (role == 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i) G[XG_V1 + i] = v1[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) G[XG_V2 + i] = v2[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            G[XG_CT + i] = ct0[i];
            G[XG_CT + 8 + i] = ct1[i];
        }
        G[XG_TH] = bad_reflector ? nan("") : th0;
        G[XG_TH + 1] = th1;
        G[XG_TH + 2] = jc;
        G[XG_TH + 3] = js;
    }
    // own slot: role A factorises (s0 I - B) and solves with c~0, role B (s1 I - B) with c~1 -- the same instructions
    // on different data; B adds the certificate at shift 0 when both slots are positive (second trip of the rolled loop)
    const bool pos0 = th0 > 0.0, pos1 = th1 > 0.0;
    const double shift_own = role ? (pos1 ? th1 : 0.0) : (pos0 ? th0 : 0.0);
    double rhs[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rhs[i] = role ? ct1[i] : ct0[i];
    bool ok = true, ok_cert = true;
    const int n_fact = (role == 1 && pos0 && pos1) ? 2 : 1;
#pragma unroll 1
    for (int q = 0; q < n_fact; ++q) {
        double xq[8];
        const bool okq = chol8_solve_reg(m, q == 0 ? shift_own : 0.0, rhs, xq);
        if (q == 0) {
            ok = okq;
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = xq[i];
        } else {
            ok_cert = okq;
        }
    }
    if (!ok) {   // (a failed certificate at shift 0 leaves the slot's own correction in place, like track_step)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) G[XG_X + 8 * role + i] = x[i];
    X[X2_OK + role] = (ok && ok_cert) ? 1.0 : 0.0;
}

// Second half (steps 4-5): Rayleigh-Ritz on span{[x0; e~0], [x1; e~1]} (both roles, same arithmetic), then each role
// assembles its own vector and takes it back through the reflectors.  Returns TRK_OK / TRK_NEED_FULL (the same in both
// roles); corr2 as in track_step.
template <int S>
CVX_HD int t2_step_p2(int role, Arr<S> U, Arr<S> TH, Arr<S> G, Arr<S> X, double& corr2)
{
    double x0[8], x1[8], ct0[8], ct1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x0[i] = G[XG_X + i];
        x1[i] = G[XG_X + 8 + i];
        ct0[i] = G[XG_CT + i];
        ct1[i] = G[XG_CT + 8 + i];
    }
    const double th0 = G[XG_TH], th1 = G[XG_TH + 1], jc = G[XG_TH + 2], js = G[XG_TH + 3];
    const bool ok = X[X2_OK] != 0.0 && X[X2_OK + 1] != 0.0;
    if (!isfinite(th0)) {   // (role A marks unusable reflectors with a NaN)
        corr2 = 0.0;
        return TRK_NEED_FULL;
    }
    const bool pos0 = th0 > 0.0, pos1 = th1 > 0.0;
    const double s0 = pos0 ? th0 : 0.0, s1 = pos1 ? th1 : 0.0;
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
    const double h00 = fma(s0, x00, c00) + th0, h11 = fma(s1, x11, c11) + th1;
    const double h01 = 0.5 * (fma(s1, x01, c01) + fma(s0, x01, c10));
    const double g00 = 1.0 + x00, g01 = x01, g11 = 1.0 + x11;
    const double i00 = cvx_rsqrt(g00);
    const double l10 = g01 * i00;
    const double i11 = cvx_rsqrt(g11 - l10 * l10);
    const double y00 = h00 * i00, y10 = (h01 - l10 * y00) * i11, y01 = h01 * i00, y11 = (h11 - l10 * y01) * i11;
    const double hp00 = y00 * i00, hp10 = y10 * i00, hp11 = (y11 - y10 * l10 * i00) * i11;
    double rc_, rs_, rt_;
    jacobi_cs(hp00, hp11, hp10, rc_, rs_, rt_);
    const double nth0 = fma(-rt_, hp10, hp00), nth1 = fma(rt_, hp10, hp11);
    if (!isfinite(nth0) || !isfinite(nth1)) return TRK_NEED_FULL;
    // f0 = a00 w0 + a01 w1,  f1 = a10 w0 + a11 w1,  w_i = [x_i; e~_i]
    const double ca = (role == 0) ? rc_ * i00 + rs_ * i11 * l10 * i00 : rs_ * i00 - rc_ * i11 * l10 * i00;
    const double cb = (role == 0) ? -rs_ * i11 : rc_ * i11;
    double f[10];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fma(ca, x0[i], cb * x1[i]);
    f[8] = fma(ca, jc, cb * js);
    f[9] = fma(ca, -js, cb * jc);
    double v1[10], v2[10], n1 = 0.0, n2 = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        v1[i] = G[XG_V1 + i];
        n1 = fma(v1[i], v1[i], n1);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        v2[i] = G[XG_V2 + i];
        n2 = fma(v2[i], v2[i], n2);
    }
    v2[9] = 0.0;
    reflect_vec<9>(f, v2, 2.0 / n2);
    reflect_vec<10>(f, v1, 2.0 / n1);
#pragma unroll
    for (int i = 0; i < 10; ++i) U[10 * role + i] = f[i];
    TH[role] = (role == 0) ? nth0 : nth1;
    return ok ? TRK_OK : TRK_NEED_FULL;
}

// What follows the eigenpair step (track_pass_eig without the step): certificate bookkeeping, penalty rescale on the
// role's share of M and Q/rho, polishing, dual objective.  Same return values as track_pass_eig.  Reads U, TH of both
// roles: call after the barrier that follows t2_step_p2.
template <int S>
CVX_HD int t2_after_step(int role, const Opts& o, Arr<S> M, Arr<S> U, Arr<S> TH, Arr<S> QR, LaneState& st, int rc, double corr2)
{
    if (rc != TRK_OK) {
        if (!st.iterating || st.bad >= CVX_TRK_TOLERATE || !isfinite(corr2)) return -1;
        ++st.bad;
    } else {
        st.bad = 0;
    }
    if (st.iterating) {
        const double rf = rescale_factor(st.it);
        if (rf > 0.0) {
            const double ic = 1.0 / rf;
            const double t0 = fmax(TH[0], 0.0), t1 = fmax(TH[1], 0.0);
            double a[10], bb[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                a[i] = U[i];
                bb[i] = U[10 + i];
            }
#pragma unroll
            for (int r = 0; r < 10; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) {
                    const int e = sidx(r, c);
                    if ((e < 28) != (role == 0)) continue;
                    const double z = fma(t0 * a[r], a[c], t1 * bb[r] * bb[c]);
                    M[e] = fma(ic, (double)M[e] - z, z);
                }
#pragma unroll 1
            for (int e = (role == 0 ? 0 : 23); e < (role == 0 ? 23 : 45); ++e) QR[e] = QR[e] * ic;
            st.rho *= rf;
            aa_reset(st.aa);
            st.res_prev = 1e300;
        }
        return 0;
    }
    if (corr2 > 1e-28 && st.phase < 8) {
        st.phase += 2;
        return 0;
    }
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

// A new problem, loaded by the two roles together: A takes M and the tracked pairs, B takes Q/rho; both set the state.
template <int ROLE, int S>
CVX_HD void t2_begin(const double* rec, const Opts& o, Arr<S> M, Arr<S> U, Arr<S> TH, Arr<S> QR, LaneState& st)
{
    const double rho = rec[TR_RHO];
    if (ROLE == 0) {
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
    } else {
        double q[45];
#pragma unroll
        for (int e = 0; e < 45; ++e) q[e] = rec[TR_Q + e];
#pragma unroll
        for (int e = 0; e < 45; ++e) QR[e] = q[e];
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

// ... from the FP32 first phase (see track_begin_warm)
template <int ROLE, int S>
CVX_HD void t2_begin_warm(const double* rec, const double* w, const Opts& o, Arr<S> M, Arr<S> U, Arr<S> TH, Arr<S> QR,
                          LaneState& st)
{
    const double rho = rec[TR_RHO];
    if (ROLE == 0) {
#pragma unroll 5
        for (int e = 0; e < 55; ++e) M[e] = w[e];
        double u[20], th[2];
        track_from_full(Arr<1>{const_cast<double*>(w) + 55}, Arr<1>{const_cast<double*>(w) + 155}, u, th);
#pragma unroll 4
        for (int e = 0; e < 20; ++e) U[e] = u[e];
        TH[0] = th[0];
        TH[1] = th[1];
    } else {
#pragma unroll 5
        for (int e = 0; e < 45; ++e) QR[e] = rec[TR_Q + e];
    }
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

// The role's half of a DR iteration (see track_pass_dr): the partial squared residual goes to X[X2_RES + ROLE].
template <int ROLE, int S>
CVX_HD void t2_dr_half(const Opts& o, Arr<S> M, Arr<S> G, Arr<S> U, Arr<S> TH, Arr<S> QR, Arr<S> X)
{
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
    X[X2_RES + ROLE] = dr_affine_part<ROLE>(M, G, QR, o.alpha, 1.0 / o.sigma, o.rowk, zf);
}

// One pass of a problem by one of its two threads: DR iteration, Anderson step, eigenpair step, bookkeeping.
// `live`: this thread pair owns a problem (in the kernel every lane walks through the barriers, live or not).
// `sync` is a barrier that includes the two threads of the pair, `vote` an OR-vote (with the same barrier semantics)
// that comes out the same in both: in the kernel both span the CTA -- the eight warps then stay in lock-step through
// the whole pass and the four warps of a role share the instruction-cache lines of the role's code (with a barrier per
// warp pair the pairs drift apart within a pass: ncu stall_no_instruction 22 %).
// Returns 1 when the problem is done (park it), 0 to continue, -1 to hand it back.
template <int S, class Hist, class Sync, class Vote>
CVX_HD int t2_pass(int role, const Opts& o, bool live, Arr<S> M, Arr<S> G, Arr<S> U, Arr<S> TH, Arr<S> QR, Arr<S> X,
                   ArrT<S, float> XF, const Hist& H, LaneState& st, int wslot, const Sync& sync, const Vote& vote)
{
    // ---- 1. DR iteration ---------------------------------------------------------------------------------------
    const bool iter = live && st.finite && st.iterating;
    if (iter) {
        if (role == 0) t2_dr_half<0>(o, M, G, U, TH, QR, X);
        else t2_dr_half<1>(o, M, G, U, TH, QR, X);
    }
    sync();
    bool want = false;
    if (iter) want = track_decide(o, st, (double)X[X2_RES] + (double)X[X2_RES + 1], M, G, role ? 28 : 0, role ? 55 : 28);
    // ---- 2. Anderson step (all threads of a warp, like aa_step) ----------------------------------------------------
    if (vote(want)) {
        float sc, nd[AA_M], ndd, ndg;
        const bool close = want && st.aa.have_prev;
        aa2_part_a(role, G, H, st.aa, want, wslot, (float)st.res_prev, sc, nd, ndd, ndg);
        st.aa.mask = close ? (st.aa.mask | (1u << wslot)) : (st.aa.mask & ~(1u << wslot));
        if (!want) st.aa.mask = 0u;
        if (role == 1) {
#pragma unroll
            for (int j = 0; j < AA_M; ++j) XF[XF_PART + j] = nd[j];
            XF[XF_PART + 7] = ndd;
            XF[XF_PART + 8] = ndg;
        }
        sync();
        uint32_t cf[9];
        if (role == 0) {
#pragma unroll
            for (int j = 0; j < AA_M; ++j) nd[j] += XF[XF_PART + j];
            ndd += XF[XF_PART + 7];
            ndg += XF[XF_PART + 8];
            aa2_solve(H, st.aa.mask, want, wslot, sc, nd, ndd, ndg, cf);
#pragma unroll
            for (int k = 0; k < 9; ++k) XF[XF_COEF + k] = w2f(cf[k]);
        }
        sync();
        if (role == 1) {
#pragma unroll
            for (int k = 0; k < 9; ++k) cf[k] = f2w(XF[XF_COEF + k]);
        }
        float ng, ns;
        aa2_part_b(role, M, G, H, want, cf, sc, ng, ns);
        XF[XF_NORM + 2 * role] = ng;
        XF[XF_NORM + 2 * role + 1] = ns;
        sync();
        const float ngt = XF[XF_NORM] + XF[XF_NORM + 2], nst = XF[XF_NORM + 1] + XF[XF_NORM + 3];
        const bool ok = cf[8] != 0u;
        const bool reject = ok && !(nst <= AA_MAX_STEP2 * ngt);
        if (H.any(reject)) aa2_part_c(role, M, H, reject, cf, sc);
        if (want && st.aa.mask != 0u && (!ok || reject)) st.aa.mask = 0u;
        st.aa.have_prev = want;
        st.aa.scale_prev = sc;
    }
    sync();
    // ---- 3. eigenpair step: up to four rounds (see track_pass_eig), uniform over the pair --------------------------------
    int rc = TRK_OK;
    double corr2 = 0.0;
    bool again = live && st.finite;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        if (k > 0 && !vote(again)) break;   // (the first round is every pass's: no vote)
        if (again) t2_step_p1(role, M, U, G, X);
        sync();
        if (again) rc = t2_step_p2(role, U, TH, G, X, corr2);
        sync();
        again = again && (rc != TRK_OK && corr2 > CVX_TRK_REPEAT2);
    }
    if (!live) return 0;
    if (!st.finite) return 1;
    return t2_after_step(role, o, M, U, TH, QR, st, rc, corr2);
}

}  // namespace cvx
