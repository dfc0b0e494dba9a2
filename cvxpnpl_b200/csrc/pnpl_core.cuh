// pnpl_core.cuh -- per-problem device routines of the CvxPnPL hot path.
//
// One *thread* owns one pose problem end to end (assembly -> 10x10 SDP by
// Douglas-Rachford/ADMM -> pose extraction).  Every per-problem array lives in a
// strided view (`Arr<S>`): element e of the problem owned by thread t sits at
// base[e * S + t], so a warp touching "element e" reads 32 consecutive doubles
// (bank-conflict free in shared memory, coalesced in global memory).
//
// Reference path being replaced (cvxpnpl.py, commit e20cca87):
//   assembly      _point_constraints 20-104, _line_constraints 107-153,
//                 B/A 548-549 | 579-580 | 623-624, Q 475
//   SDP solve     scs.solve 485-489 on the static data of 387-448
//   extraction    493-520, _constraint_ortho_det 221-343, _re6q3 156-218
//
// The routines are written as __host__ __device__ so the very same code can be
// compiled by the host compiler inside tests/ (tests/host_harness.cpp) to debug
// numerics without a GPU.  The product library only ever launches them as CUDA
// kernels (pnpl_kernels.cu); there is no CPU execution path in the product.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CVX_HD __host__ __device__ __forceinline__
#define CVX_HD_NOINLINE __host__ __device__ __noinline__
// compiler-only scheduling fence: keeps ptxas from hoisting every load of an
// unrolled region to its top (which blows the register budget)
#define CVX_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define CVX_SCHED_FENCE() asm volatile("" ::: "memory")
#define CVX_HD inline
#define CVX_HD_NOINLINE
#endif

namespace cvx {

// ---- status codes written per problem (low byte) -------------------------------
enum : int32_t {
    ST_OK = 0,            // converged to eps
    ST_MAX_ITERS = 1,     // iteration cap hit: last iterate returned (SCS "solved_inaccurate")
    ST_NAN = 2,           // non-finite data/iterate -> single NaN pose (cvxpnpl.py:493-498)
    ST_SINGULAR = 3,      // singular system in the rank-4 recovery (LinAlgError at cvxpnpl.py:165/212)
    ST_RANK0 = 4,         // no eigenvalue above 1e-3 (NotImplementedError at cvxpnpl.py:341)
    ST_FLAG_NOT_CERTIFIED = 0x100  // |obj - dual obj| > eps (the warning of cvxpnpl.py:516-519)
};

template <int S, class R>
struct ArrT {
    R* p;
    CVX_HD R& operator[](int e) const { return p[(size_t)e * S]; }
    CVX_HD ArrT<S, R> sub(int off) const { return ArrT<S, R>{p + (size_t)off * S}; }
};
template <int S>
using Arr = ArrT<S, double>;

// runtime-strided read-only / write view for global scratch
struct GArr {
    double* p;
    int64_t stride;
    CVX_HD double& operator[](int e) const { return p[(int64_t)e * stride]; }
};

template <class R>
struct GArrT {
    R* p;
    int64_t stride;
    CVX_HD R& operator[](int e) const { return p[(int64_t)e * stride]; }
};

// packed lower-triangular (row-major) index of a symmetric matrix
CVX_HD int sidx(int i, int j) { return i >= j ? (i * (i + 1)) / 2 + j : (j * (j + 1)) / 2 + i; }

// ---- the 15 "triple" equalities (rows 2,3,5,8,9,11,13..21 of the reference's
// A, cvxpnpl.py:401-435): sum_k s_k Z[i_k, j_k] = 0 over off-diagonal entries.
// Each off-diagonal entry of Z belongs to exactly one triple.  (i > j always.)
// The last macro argument marks the three ROW-orthogonality triples, which (with the
// three row-norm equalities) the "rc" ablation of benchmarks/toolkit/methods/rc.py drops.
// (two halves: the two-threads-per-problem form of the tracked solver gives the first eight to one thread, the last
// seven and the diagonal equalities to the other)
#define CVX_TRIPLES_A(X)                                      \
    X(1, 0, +1, 4, 3, +1, 7, 6, +1, 1)   /* r0.r1 */          \
    X(2, 0, +1, 5, 3, +1, 8, 6, +1, 1)   /* r0.r2 */          \
    X(2, 1, +1, 5, 4, +1, 8, 7, +1, 1)   /* r1.r2 */          \
    X(3, 0, +1, 4, 1, +1, 5, 2, +1, 0)   /* c0.c1 */          \
    X(6, 0, +1, 7, 1, +1, 8, 2, +1, 0)   /* c0.c2 */          \
    X(6, 3, +1, 7, 4, +1, 8, 5, +1, 0)   /* c1.c2 */          \
    X(5, 1, +1, 4, 2, -1, 9, 6, -1, 0)   /* (c0xc1)_0 = c2_0 */ \
    X(3, 2, +1, 5, 0, -1, 9, 7, -1, 0)
#define CVX_TRIPLES_B(X)                                      \
    X(4, 0, +1, 3, 1, -1, 9, 8, -1, 0)                        \
    X(8, 4, +1, 7, 5, -1, 9, 0, -1, 0)   /* c1xc2 = c0 */     \
    X(6, 5, +1, 8, 3, -1, 9, 1, -1, 0)                        \
    X(7, 3, +1, 6, 4, -1, 9, 2, -1, 0)                        \
    X(7, 2, +1, 8, 1, -1, 9, 3, -1, 0)   /* c2xc0 = c1 */     \
    X(8, 0, +1, 6, 2, -1, 9, 4, -1, 0)                        \
    X(6, 1, +1, 7, 0, -1, 9, 5, -1, 0)
#define CVX_TRIPLES(X) CVX_TRIPLES_A(X) CVX_TRIPLES_B(X)

// ---------------------------------------------------------------------------------
// Assembly: correspondences -> Q (45 unique entries of the 9x9 block, packed lower)
// and B (3x9, row-major).  Uses the Kronecker structure of the reference's rows:
// every correspondence contributes  C'C += (P P') (x) W,  N'C += P' (x) W,
// N'N += W  with the 3x3 weight  W = [p]x'[p]x = |p|^2 I - p p'  for a point with
// bearing p (cvxpnpl.py:37, 53-102) and  W = n n'  for a line endpoint with
// plane normal n (cvxpnpl.py:123-153).  Then B = (N'N)^-1 N'C and
// Q = A'A = C'C - (N'C)' B   (cvxpnpl.py:623-624, 475; A never materialised).
// ---------------------------------------------------------------------------------
struct Accum {
    double PPW[6][6];  // [sym idx of (a,c)][sym idx of (i,j)]  -> C'C
    double PW[3][6];   // [a][sym idx (i,j)]                     -> N'C
    double W[6];       // sym idx (i,j)                          -> N'N
};

CVX_HD int sym3(int i, int j) { return i >= j ? (i * (i + 1)) / 2 + j : (j * (j + 1)) / 2 + i; }

CVX_HD void accum_init(Accum& a)
{
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        a.W[i] = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) a.PPW[i][j] = 0;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) a.PW[i][j] = 0;
}

CVX_HD void accum_add(Accum& a, const double P[3], const double W[6])
{
    double PP[6] = {P[0] * P[0], P[1] * P[0], P[1] * P[1], P[2] * P[0], P[2] * P[1], P[2] * P[2]};
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        a.W[e] += W[e];
#pragma unroll
        for (int g = 0; g < 6; ++g) a.PPW[g][e] = fma(PP[g], W[e], a.PPW[g][e]);
#pragma unroll
        for (int g = 0; g < 3; ++g) a.PW[g][e] = fma(P[g], W[e], a.PW[g][e]);
    }
}

// general 3x3 inverse (adjugate / determinant); K is row-major
CVX_HD void inv3(const double K[9], double Ki[9])
{
    double c00 = K[4] * K[8] - K[5] * K[7];
    double c01 = K[5] * K[6] - K[3] * K[8];
    double c02 = K[3] * K[7] - K[4] * K[6];
    double det = K[0] * c00 + K[1] * c01 + K[2] * c02;
    double id = 1.0 / det;
    Ki[0] = c00 * id;
    Ki[1] = (K[2] * K[7] - K[1] * K[8]) * id;
    Ki[2] = (K[1] * K[5] - K[2] * K[4]) * id;
    Ki[3] = c01 * id;
    Ki[4] = (K[0] * K[8] - K[2] * K[6]) * id;
    Ki[5] = (K[2] * K[3] - K[0] * K[5]) * id;
    Ki[6] = c02 * id;
    Ki[7] = (K[1] * K[6] - K[0] * K[7]) * id;
    Ki[8] = (K[0] * K[4] - K[1] * K[3]) * id;
}

CVX_HD void bearing(const double Ki[9], double u, double v, double p[3])
{
    p[0] = fma(Ki[0], u, fma(Ki[1], v, Ki[2]));
    p[1] = fma(Ki[3], u, fma(Ki[4], v, Ki[5]));
    p[2] = fma(Ki[6], u, fma(Ki[7], v, Ki[8]));
}

// From the accumulated sums to B = (N'N)^-1 N'C and Q = C'C - (N'C)' B.
template <class QOut, class BOut>
CVX_HD bool reduce_accum(const Accum& acc, QOut Q, BOut Bm)
{
    // inverse of the symmetric 3x3 N'N
    double G[9] = {acc.W[0], acc.W[1], acc.W[3], acc.W[1], acc.W[2], acc.W[4], acc.W[3], acc.W[4], acc.W[5]};
    double Gi[9];
    inv3(G, Gi);
    // N'C [i][3a+j] = PW[a][sym(i,j)];  B = Gi * N'C
    double Bl[27];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s = fma(Gi[3 * i + k], acc.PW[a][sym3(k, j)], s);
                Bl[9 * i + 3 * a + j] = s;
            }
    // Q[(3a+i),(3c+j)] = PPW[sym(a,c)][sym(i,j)] - sum_k N'C[k][3a+i] * B[k][3c+j]
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 9; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            const int a = r / 3, i = r % 3, cc = c / 3, j = c % 3;
            double s = acc.PPW[sym3(a, cc)][sym3(i, j)];
#pragma unroll
            for (int k = 0; k < 3; ++k) s = fma(-acc.PW[a][sym3(k, i)], Bl[9 * k + 3 * cc + j], s);
            Q[sidx(r, c)] = s;
            ok = ok && isfinite(s);
        }
#pragma unroll
    for (int i = 0; i < 27; ++i) {
        Bm[i] = Bl[i];
        ok = ok && isfinite(Bl[i]);
    }
    return ok;
}

// Q: 45 packed (9x9 lower, row-major), Bm: 27 (3x9 row-major).  Generic output
// accessors so the caller can target registers, shared or global memory.
// Returns false if the 3x3 normal system is singular / data non-finite.
template <class QOut, class BOut>
CVX_HD bool assemble(const double* K, const double* pts_2d, const double* pts_3d, int n_pts,
                     const double* line_2d, const double* line_3d, int n_lines, QOut Q, BOut Bm)
{
    double Kl[9], Ki[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Kl[i] = K[i];
    inv3(Kl, Ki);

    Accum acc;
    accum_init(acc);
    for (int i = 0; i < n_pts; ++i) {
        double p[3], P[3] = {pts_3d[3 * i], pts_3d[3 * i + 1], pts_3d[3 * i + 2]};
        bearing(Ki, pts_2d[2 * i], pts_2d[2 * i + 1], p);
        double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        double W[6] = {n2 - p[0] * p[0], -p[1] * p[0], n2 - p[1] * p[1],
                       -p[2] * p[0],     -p[2] * p[1], n2 - p[2] * p[2]};
        accum_add(acc, P, W);
    }
    for (int i = 0; i < n_lines; ++i) {
        double a[3], b[3];
        bearing(Ki, line_2d[4 * i], line_2d[4 * i + 1], a);
        bearing(Ki, line_2d[4 * i + 2], line_2d[4 * i + 3], b);
        double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        double inv = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        n[0] *= inv; n[1] *= inv; n[2] *= inv;
        double W[6] = {n[0] * n[0], n[1] * n[0], n[1] * n[1], n[2] * n[0], n[2] * n[1], n[2] * n[2]};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double P[3] = {line_3d[6 * i + 3 * e], line_3d[6 * i + 3 * e + 1], line_3d[6 * i + 3 * e + 2]};
            accum_add(acc, P, W);
        }
    }
    return reduce_accum(acc, Q, Bm);
}

// ---- type-generic strided view ---------------------------------------------------
CVX_HD constexpr int jp_p(int k) { return k; }          // fixed pairs (k, 9-k), k = 0..4
CVX_HD constexpr int jp_q(int k) { return 9 - k; }
CVX_HD constexpr int jp_sigma(int i) { return i == 0 ? 0 : (i == 9 ? 1 : i + 1); }


}  // namespace cvx

// CVX_RELSKIP: pivots smaller than this relative to the diagonal gap are not rotated;
// CVX_TINY: guard against d^2 + b^2 underflowing.
#define CVX_REAL double
#define CVX_RELSKIP 1e-18
#define CVX_TINY 1e-280
namespace cvx {
#include "pnpl_dr.inl"
}  // namespace cvx
#undef CVX_REAL
#undef CVX_RELSKIP
#undef CVX_TINY
#define CVX_REAL float
#define CVX_RELSKIP 1e-9f
#define CVX_TINY 1e-36f
namespace cvx {
namespace f32 {
#include "pnpl_dr.inl"
}  // namespace f32
}  // namespace cvx
#undef CVX_REAL
#undef CVX_RELSKIP
#undef CVX_TINY

namespace cvx {

// Rotation with an FP32 angle and an FP64 normalisation: tan(theta) from the textbook
// formula in single precision (one MUFU square root and one reciprocal), then
// c = 1 / sqrt(1 + t^2), s = t c in double.  The rotation is orthogonal to double precision
// (which is what keeps V orthonormal); the pivot is annihilated only to ~1e-7 of its size
// instead of exactly, which a warm-started solver that sweeps once per DR iteration does
// not notice (the leftover is squared away by the next sweep).  Half the FP64 work of
// jacobi_cs (one reciprocal square root instead of two): used where the rotation angles
// are on the critical path (the warp-per-problem kernel, pnpl_warp.cuh).
CVX_HD void jacobi_cs_fast(double app, double aqq, double apq, double& c, double& s)
{
    const double d = aqq - app, b2 = 2.0 * apq;
    const double g = fma(d, d, b2 * b2);
    const bool skip = !(fabs(apq) > 1e-18 * fabs(d)) || !(g > 1e-280);
    const float df = (float)d, bf = (float)b2;
#if defined(__CUDA_ARCH__)
    // device: MUFU reciprocal square root / reciprocal without the IEEE slow paths (the five chains of a round are the
    // critical path of a warp's sweep: ~50 of ~190 instructions per round).  h = 0 (both |d| and |b| below 1e-19)
    // gives NaN -> identity rotation: such a pivot is far below what the sweep resolves anyway.
    const float h = fmaf(df, df, bf * bf);
    float tf = __fdividef(copysignf(bf, bf * df), fabsf(df) + h * rsqrtf(h));
#else
    float tf = copysignf(bf, bf * df) / (fabsf(df) + sqrtf(fmaf(df, df, bf * bf)));
#endif
    tf = fminf(fmaxf(tf, -1.f), 1.f);        // |theta| <= pi/4; also catches 0/0-free overflow to inf
    const double t = skip || !(tf == tf) ? 0.0 : (double)tf;
    c = cvx_rsqrt(fma(t, t, 1.0));
    s = t * c;
}

// ---------------------------------------------------------------------------------
// Anderson acceleration (type II, memory AA_M = 7) of the DR fixed-point iteration
// M <- F(M), the accelerator SCS itself relies on.  With g_k = F(M_k) - M_k:
//     gamma = argmin || g_k - dG gamma ||,   M_{k+1} = M_k + g_k - (dM + dG) gamma
// where the columns of dG / dM are the last AA_M differences of g and M.  Memory
// matters: measured median DR iterations on PnPL 8+4 are 203 (none), 101 (3), 80 (5),
// 70 (7), 62 (10).  Seven columns is what one thread's share of tensor memory (512
// 32-bit words) holds once the columns are stored as FP16 pairs:
//     word   0.. 55   g_{k-1}                                   FP32
//     word  56.. 83   step_{k-1} * s_{k-1}                      FP16 x 2
//     word  84..279   dG_j * s_j      [chunk c][column j][4]    FP16 x 2
//     word 280..475   (dM_j+dG_j)*s_j [chunk c][column j][4]    FP16 x 2
//     word 476..503   Gram matrix of the stored dG columns      FP32 (packed lower)
// Every column pair (dG_j, dM_j + dG_j) is multiplied by a power of two s_j that
// brings |g| at the time of writing to ~16 before the FP16 rounding, so FP16's 11-bit
// mantissa is spent on the entries that matter whatever the residual is (1e-2 ..
// 1e-9).  The scale never has to be stored: scaling both halves of a pair by s_j only
// rescales gamma_j, the extrapolated point is unchanged.  (Plain bf16 columns were
// measured too: they cost the hard problems -- PnL 6, 4 points -- 15-20 % more
// iterations; scaled FP16 is within 2 % of FP32 columns.)  The history only steers
// the extrapolation; the fixed point, and hence the result, does not depend on it.
// Inner products are plain Euclidean on the packed 55-vector.  Safeguard: an
// extrapolated step longer than 10 |g_k| (or a Gram matrix that is not positive
// definite) is rejected, the history dropped and the plain step kept.
// On entry M already holds M_k + g_k and G holds g_k; on exit M holds M_{k+1}.
// ---------------------------------------------------------------------------------
constexpr int AA_M = 7;
#ifndef CVX_AA_ON
#define CVX_AA_ON 0.4
#endif
// accelerate only once ||X - Z||_F < 0.4 (|Z| ~ 4).  Was 0.15 while an extrapolation during the early
// active-set changes could strand M on a plateau; with the plateau jump (pnpl_solve.cuh) a higher
// threshold is safe and saves ~2 % of the iterations (host build, 3000 problems each: PnPL 8+4 mean
// 55.5 -> 54.5, PnP-8 56.8 -> 55.6, PnL-6 78.2 -> 75.1; 0.25 ... 1.0 are all within 0.5 %)
constexpr double AA_RES2_ON = CVX_AA_ON * CVX_AA_ON;
// the FP32 first phase hands a problem to the FP64 solver at ||X - Z||_F < 0.15 (FP32 residual floor ~1e-5)
constexpr double FP32_EXIT_RES2 = 0.15 * 0.15;
#ifndef CVX_AA_MAXSTEP2
#define CVX_AA_MAXSTEP2 100.f
#endif
constexpr float AA_MAX_STEP2 = CVX_AA_MAXSTEP2;   // an extrapolated step longer than 10 |g| is rejected
constexpr int AA_CHUNKS = 7;                  // 7 chunks x 8 entries >= 55
constexpr int AA_OFF_GP = 0;
constexpr int AA_OFF_SP = 56;
constexpr int AA_OFF_DG = 84;
constexpr int AA_OFF_DS = AA_OFF_DG + AA_CHUNKS * AA_M * 4;   // 280
constexpr int AA_OFF_GRAM = AA_OFF_DS + AA_CHUNKS * AA_M * 4; // 476
constexpr int AA_GRAM_WORDS = AA_M * (AA_M + 1) / 2;          // 28 (loaded as a 32-word block)
constexpr int AA_WORDS = 512;                                 // = all TMEM columns of a lane
static_assert(AA_OFF_GRAM + 32 <= AA_WORDS, "history does not fit one TMEM lane");

struct AAState {
    uint32_t mask;      // valid history columns (bit j = column j)
    bool have_prev;     // g_prev / step_prev hold the previous iteration of THIS problem
    float scale_prev;   // power of two the stored step_prev was multiplied by
};

CVX_HD void aa_reset(AAState& aa)
{
    aa.mask = 0u;
    aa.have_prev = false;
    aa.scale_prev = 1.f;
}

// ---- FP16 pair <-> two floats, 32-bit word <-> float --------------------------------
CVX_HD uint32_t f2w(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
#endif
}
CVX_HD float w2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x;
    memcpy(&x, &u, 4);
    return x;
#endif
}
CVX_HD uint32_t pack_h2(float lo, float hi)
{
#if defined(__CUDA_ARCH__)
    uint32_t w;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    return w;
#else
    const _Float16 a = (_Float16)lo, b = (_Float16)hi;
    uint16_t ua, ub;
    memcpy(&ua, &a, 2);
    memcpy(&ub, &b, 2);
    return (uint32_t)ua | ((uint32_t)ub << 16);
#endif
}
CVX_HD void unpack_h2(uint32_t w, float& lo, float& hi)
{
#if defined(__CUDA_ARCH__)
    asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(w));
#else
    const uint16_t ua = (uint16_t)(w & 0xffffu), ub = (uint16_t)(w >> 16);
    _Float16 a, b;
    memcpy(&a, &ua, 2);
    memcpy(&b, &ub, 2);
    lo = (float)a;
    hi = (float)b;
#endif
}

// Mixed-precision FMA  acc + half(a) * half(b)  in FP32: one instruction on sm_100a (SASS FHFMA, the operand
// halves are selected inside the instruction), so a dot product with an FP16-pair history column needs no
// unpacking.  The product of two FP16 numbers is exact in FP32, so host and device agree bit for bit.
CVX_HD float hfma_ll(uint32_t a, uint32_t b, float acc)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("{.reg .f16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.f16 %0, al, bl, %3;}"
        : "=f"(r) : "r"(a), "r"(b), "f"(acc));
    return r;
#else
    float al, ah, bl, bh;
    unpack_h2(a, al, ah);
    unpack_h2(b, bl, bh);
    return fmaf(al, bl, acc);
#endif
}
CVX_HD float hfma_hh(uint32_t a, uint32_t b, float acc)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("{.reg .f16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.f16 %0, ah, bh, %3;}"
        : "=f"(r) : "r"(a), "r"(b), "f"(acc));
    return r;
#else
    float al, ah, bl, bh;
    unpack_h2(a, al, ah);
    unpack_h2(b, bl, bh);
    return fmaf(ah, bh, acc);
#endif
}
CVX_HD float hfma_hl(uint32_t a, uint32_t b, float acc)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("{.reg .f16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.f16 %0, ah, bl, %3;}"
        : "=f"(r) : "r"(a), "r"(b), "f"(acc));
    return r;
#else
    float al, ah, bl, bh;
    unpack_h2(a, al, ah);
    unpack_h2(b, bl, bh);
    return fmaf(ah, bl, acc);
#endif
}

// power of two s with |g| s in [16, 64) for |g|^2 ~ res2 (the squared DR residual)
CVX_HD float aa_scale(float res2)
{
    const int e = (int)((f2w(res2) >> 23) & 0xffu) - 127;
    int k = 4 - (e >> 1);
    k = k > 100 ? 100 : (k < -100 ? -100 : k);
    return w2f((uint32_t)(k + 127) << 23);
}

// History accessors.  Both expose N-word loads/stores at a word offset of the
// thread's history (N = 4, 8, 16, 32):
//   * HistMem  -- plain strided memory (host harness; the stage kernel's global scratch)
//   * HistTmem -- Blackwell tensor memory (pnpl_kernels.cu): every thread owns one TMEM
//                 lane = 512 private 32-bit words, 12-cycle loads, no shared-memory or
//                 L2 traffic.  tcgen05.ld/st are warp-collective with a warp-uniform
//                 address, which is why aa_step below is written so that ALL lanes of a
//                 warp execute every history access (column slots are warp-uniform,
//                 lanes that do not accelerate just compute on dead data).
struct HistMem {
    uint32_t* p;
    int64_t stride;
    template <int N>
    CVX_HD void ld(int off, uint32_t* o) const
    {
#pragma unroll
        for (int u = 0; u < N; ++u) o[u] = p[(int64_t)(off + u) * stride];
    }
    template <int N>
    CVX_HD void st(int off, const uint32_t* v) const
    {
#pragma unroll
        for (int u = 0; u < N; ++u) p[(int64_t)(off + u) * stride] = v[u];
    }
    CVX_HD void wait_ld() const {}
    CVX_HD void wait_st() const {}
    CVX_HD bool any(bool f) const { return f; }
};

// all AA_M columns of chunk c (AA_M x 4 = 28 words) of the block starting at `off`
template <class Hist>
CVX_HD void aa_ld_cols(const Hist& H, int off, int c, uint32_t w[28])
{
    H.template ld<16>(off + c * 28, w);
    H.template ld<8>(off + c * 28 + 16, w + 16);
    H.template ld<4>(off + c * 28 + 24, w + 24);
}

// (round-1 form, kept for A/B runs: -DCVX_AA_V1)  One accelerated step.  `active` lanes own a problem in the tail of its DR
// iteration; `wslot` (warp-uniform, cycles 0..AA_M-1) is the column overwritten now;
// res2 is the squared DR residual of the iterate (sets the FP16 scale).  On entry M
// holds M_k + g_k and G holds g_k (G[55] is a dedicated zero so whole 8-entry chunks
// can be processed without a bounds test); on exit (active lanes) M holds M_{k+1}.
// The least squares runs on FP32 sums with an FP32 Cholesky factorisation.  The
// Gram matrix is kept across steps; only the row of the new column is recomputed.
template <int S, class RT, class Hist>
CVX_HD void aa_step_v1(ArrT<S, RT> M, ArrT<S, RT> G, const Hist& H, AAState& aa, bool active, int wslot, float res2)
{
    const bool close = active && aa.have_prev;
    // the scale may grow at most 4x per step: then (|g_{k-1}| + |g_k|) s_k and the stored
    // step (at most 10 |g|) stay far below the FP16 maximum however fast the residual drops
    const float sc = close ? fminf(aa_scale(res2), 4.f * aa.scale_prev) : aa_scale(res2);
    const float ratio = sc / aa.scale_prev;   // both powers of two
    // ---- A. close the newest column (dG = g_k - g_{k-1}, dM + dG = step_{k-1} + dG);
    //         dot products of the new column and of g with every stored column ----------
    float rg[AA_M], nd[AA_M];   // col_j . g,  dg_new . col_j  (col_wslot is the one being replaced)
    float ndd = 0.f, ndg = 0.f;  // dg_new . dg_new, dg_new . g
#pragma unroll
    for (int j = 0; j < AA_M; ++j) rg[j] = nd[j] = 0.f;
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t gpw[8], spw[4], cw[28], dgw[4], dsw[4];
        H.template ld<8>(AA_OFF_GP + 8 * c, gpw);
        H.template ld<4>(AA_OFF_SP + 4 * c, spw);
        aa_ld_cols(H, AA_OFF_DG, c, cw);
        H.wait_ld();
        float gf[8], dr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) gf[u] = (float)G[c * 8 + u];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            float s0, s1;
            unpack_h2(spw[w], s0, s1);
            const float d0 = (gf[2 * w] - w2f(gpw[2 * w])) * sc, d1 = (gf[2 * w + 1] - w2f(gpw[2 * w + 1])) * sc;
            // a column that is not closed is stored as zeros (never garbage: 0 * Inf = NaN)
            dgw[w] = close ? pack_h2(d0, d1) : 0u;
            dsw[w] = close ? pack_h2(fmaf(s0, ratio, d0), fmaf(s1, ratio, d1)) : 0u;
            unpack_h2(dgw[w], dr[2 * w], dr[2 * w + 1]);   // the rounded column is the one that counts
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            ndd = fmaf(dr[u], dr[u], ndd);
            ndg = fmaf(dr[u], gf[u], ndg);
        }
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                float c0, c1;
                unpack_h2(cw[4 * j + w], c0, c1);
                rg[j] = fmaf(c0, gf[2 * w], fmaf(c1, gf[2 * w + 1], rg[j]));
                nd[j] = fmaf(c0, dr[2 * w], fmaf(c1, dr[2 * w + 1], nd[j]));
            }
        H.template st<4>(AA_OFF_DG + c * 28 + 4 * wslot, dgw);
        H.template st<4>(AA_OFF_DS + c * 28 + 4 * wslot, dsw);
    }
    // the overwritten slot is valid only if this lane had a previous iterate
    aa.mask = close ? (aa.mask | (1u << wslot)) : (aa.mask & ~(1u << wslot));
    if (!active) aa.mask = 0u;
    // refresh row / column wslot of the Gram matrix and the right-hand side
    float gram[32];
    {
        uint32_t gw[32];
        H.template ld<32>(AA_OFF_GRAM, gw);
        H.wait_ld();
#pragma unroll
        for (int i = 0; i < AA_M; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const int e = (i * (i + 1)) / 2 + j;
                float v = w2f(gw[e]);
                if (i == wslot && j == wslot) v = ndd;
                else if (i == wslot) v = nd[j];
                else if (j == wslot) v = nd[i];
                gram[e] = v;
                gw[e] = f2w(v);
            }
        H.template st<32>(AA_OFF_GRAM, gw);
    }
#pragma unroll
    for (int j = 0; j < AA_M; ++j)
        if (j == wslot) rg[j] = ndg;
    // normal equations (FP32 Cholesky: the Gram matrix is made of FP32 sums anyway, and
    // measured iteration counts are the same as with an FP64 factorisation); columns that are
    // not valid are cut out
    float A[AA_GRAM_WORDS], r[AA_M];
    float tr = 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) {
        const bool vi = (aa.mask >> i) & 1u;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const bool vj = (aa.mask >> j) & 1u;
            A[(i * (i + 1)) / 2 + j] = (vi && vj) ? gram[(i * (i + 1)) / 2 + j] : 0.f;
        }
        r[i] = vi ? rg[i] : 0.f;
        tr += A[(i * (i + 1)) / 2 + i];
    }
    bool pd = tr > 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) A[(i * (i + 1)) / 2 + i] += 1e-6f * tr + (((aa.mask >> i) & 1u) ? 0.f : 1.f);
    // A = L L' in place (L[i][i] holds the RECIPROCAL pivot), then two triangular solves.
    // All loops have constant bounds with guards, so they unroll completely and A, r
    // stay in registers.
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
    bool ok = active && aa.mask != 0u && pd;
    float f[AA_M];
#pragma unroll
    for (int j = 0; j < AA_M; ++j) ok = ok && isfinite(r[j]);
#pragma unroll
    for (int j = 0; j < AA_M; ++j) f[j] = (ok && ((aa.mask >> j) & 1u)) ? r[j] : 0.f;
    H.wait_st();

    // ---- B. extrapolate optimistically: M -= sum_j gamma_j (dM_j + dG_j); remember g_k
    //         and the step; measure |step| against |g| ---------------------------------
    float ng = 0.f, ns = 0.f;
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t cw[28], gkw[8], spw[4];
        aa_ld_cols(H, AA_OFF_DS, c, cw);
        H.wait_ld();
        float adj[8], stp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) adj[u] = 0.f;
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                float c0, c1;
                unpack_h2(cw[4 * j + w], c0, c1);
                adj[2 * w] = fmaf(f[j], c0, adj[2 * w]);
                adj[2 * w + 1] = fmaf(f[j], c1, adj[2 * w + 1]);
            }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float gf = active ? (float)G[c * 8 + u] : 0.f;
            gkw[u] = f2w(gf);
            stp[u] = gf - adj[u];
            if (ok) M[c * 8 + u] -= (RT)adj[u];
            ng = fmaf(gf, gf, ng);
            ns = fmaf(stp[u], stp[u], ns);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) spw[w] = pack_h2(stp[2 * w] * sc, stp[2 * w + 1] * sc);
        H.template st<8>(AA_OFF_GP + 8 * c, gkw);
        H.template st<4>(AA_OFF_SP + 4 * c, spw);
    }
    H.wait_st();
    // ---- C. (rare) the extrapolated step is longer than 10 |g|: undo, keep the plain
    //         step, drop the history -----------------------------------------------------
    const bool reject = ok && !(ns <= AA_MAX_STEP2 * ng);
    if (H.any(reject)) {
#pragma unroll 1
        for (int c = 0; c < AA_CHUNKS; ++c) {
            uint32_t cw[28], gkw[8], spw[4];
            aa_ld_cols(H, AA_OFF_DS, c, cw);
            H.template ld<8>(AA_OFF_GP + 8 * c, gkw);
            H.template ld<4>(AA_OFF_SP + 4 * c, spw);
            H.wait_ld();
            float adj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) adj[u] = 0.f;
#pragma unroll
            for (int j = 0; j < AA_M; ++j)
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    float c0, c1;
                    unpack_h2(cw[4 * j + w], c0, c1);
                    adj[2 * w] = fmaf(f[j], c0, adj[2 * w]);
                    adj[2 * w + 1] = fmaf(f[j], c1, adj[2 * w + 1]);
                }
            if (reject) {
#pragma unroll
                for (int u = 0; u < 8; ++u) M[c * 8 + u] += (RT)adj[u];
#pragma unroll
                for (int w = 0; w < 4; ++w) spw[w] = pack_h2(w2f(gkw[2 * w]) * sc, w2f(gkw[2 * w + 1]) * sc);
            }
            H.template st<4>(AA_OFF_SP + 4 * c, spw);
        }
        H.wait_st();
    }
    if (active && aa.mask != 0u && (!ok || reject)) aa.mask = 0u;
    aa.have_prev = active;
    aa.scale_prev = sc;
}

// One accelerated step.  `active` lanes own a problem in the tail of its DR
// iteration; `wslot` (warp-uniform, cycles 0..AA_M-1) is the column overwritten now;
// res2 is the squared DR residual of the iterate (sets the FP16 scale).  On entry M
// holds M_k + g_k and G holds g_k (G[55] is a dedicated zero so whole 8-entry chunks
// can be processed without a bounds test); on exit (active lanes) M holds M_{k+1}.
// The least squares runs on FP32 sums with an FP32 Cholesky factorisation.  The
// Gram matrix is kept across steps; only the row of the new column is recomputed.
//
// Instruction budget (with the tracked PSD projection this step was half of the solver's
// instructions; ncu, profiles/r2c):
//   * every product with a history column is ONE mixed-precision FMA on the packed FP16 pair
//     (hfma_*: SASS FHFMA) -- no unpacking;
//   * the right-hand side  rg_j = col_j . g_k  is not recomputed: g_k = g_{k-1} + dg, and
//     col_j . dg is the Gram row nd_j that is computed anyway, so rg_j += nd_j / s (the FP16-rounded
//     dg is used, i.e. the least squares sees g up to the rounding of its last <= AA_M differences: a
//     column lives AA_M steps, then its rg is set afresh);  rg lives next to the Gram matrix;
//   * the coefficients gamma_j enter the extrapolation as FP16 numbers (common power-of-two scale), again
//     one FHFMA per entry.  The extrapolation only steers the iteration; the fixed point does not depend
//     on it.  Measured (host build, 3000 problems each): same iteration counts as the round-1 form.
constexpr int AA_OFF_RG = AA_OFF_GRAM + AA_GRAM_WORDS;   // 504: rg_0..6 (words 504..507 travel with the Gram block)
static_assert(AA_OFF_RG + 8 <= AA_WORDS, "rg does not fit");
template <int S, class RT, class Hist>
CVX_HD void aa_step(ArrT<S, RT> M, ArrT<S, RT> G, const Hist& H, AAState& aa, bool active, int wslot, float res2)
{
    const bool close = active && aa.have_prev;
    // the scale may grow at most 4x per step: then (|g_{k-1}| + |g_k|) s_k and the stored
    // step (at most 10 |g|) stay far below the FP16 maximum however fast the residual drops
    const float sc = close ? fminf(aa_scale(res2), 4.f * aa.scale_prev) : aa_scale(res2);
    const float ratio = sc / aa.scale_prev;   // both powers of two
    // ---- A. close the newest column (dG = g_k - g_{k-1}, dM + dG = step_{k-1} + dG);
    //         Gram row of the new column ---------------------------------------------------
    float nd[AA_M];              // dg_new . col_j  (col_wslot is the one being replaced)
    float ndd = 0.f, ndg = 0.f;  // dg_new . dg_new, dg_new . g
#pragma unroll
    for (int j = 0; j < AA_M; ++j) nd[j] = 0.f;
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t gpw[8], spw[4], cw[28], dgw[4], dsw[4];
        H.template ld<8>(AA_OFF_GP + 8 * c, gpw);
        H.template ld<4>(AA_OFF_SP + 4 * c, spw);
        aa_ld_cols(H, AA_OFF_DG, c, cw);
        H.wait_ld();
        float gf[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) gf[u] = (float)G[c * 8 + u];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            float s0, s1, r0, r1;
            unpack_h2(spw[w], s0, s1);
            const float d0 = (gf[2 * w] - w2f(gpw[2 * w])) * sc, d1 = (gf[2 * w + 1] - w2f(gpw[2 * w + 1])) * sc;
            // a column that is not closed is stored as zeros (never garbage: 0 * Inf = NaN)
            dgw[w] = close ? pack_h2(d0, d1) : 0u;
            dsw[w] = close ? pack_h2(fmaf(s0, ratio, d0), fmaf(s1, ratio, d1)) : 0u;
            unpack_h2(dgw[w], r0, r1);   // the rounded column is the one that counts
            ndd = fmaf(r0, r0, fmaf(r1, r1, ndd));
            ndg = fmaf(r0, gf[2 * w], fmaf(r1, gf[2 * w + 1], ndg));
        }
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 4; ++w) nd[j] = hfma_hh(cw[4 * j + w], dgw[w], hfma_ll(cw[4 * j + w], dgw[w], nd[j]));
        H.template st<4>(AA_OFF_DG + c * 28 + 4 * wslot, dgw);
        H.template st<4>(AA_OFF_DS + c * 28 + 4 * wslot, dsw);
    }
    // the overwritten slot is valid only if this lane had a previous iterate
    aa.mask = close ? (aa.mask | (1u << wslot)) : (aa.mask & ~(1u << wslot));
    if (!active) aa.mask = 0u;
    // refresh row / column wslot of the Gram matrix; right-hand side rg_j += nd_j / s (new column: dg . g)
    float gram[32], rg[AA_M];
    {
        uint32_t gw[32], rw[4];
        H.template ld<32>(AA_OFF_GRAM, gw);
        H.template ld<4>(AA_OFF_RG + 4, rw);
        H.wait_ld();
#pragma unroll
        for (int i = 0; i < AA_M; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const int e = (i * (i + 1)) / 2 + j;
                float v = w2f(gw[e]);
                if (i == wslot && j == wslot) v = ndd;
                else if (i == wslot) v = nd[j];
                else if (j == wslot) v = nd[i];
                gram[e] = v;
                gw[e] = f2w(v);
            }
        const float isc = 1.f / sc;
#pragma unroll
        for (int j = 0; j < AA_M; ++j) {
            const float prev = w2f(j < 4 ? gw[AA_GRAM_WORDS + j] : rw[j - 4]);
            rg[j] = (j == wslot) ? ndg : fmaf(nd[j], isc, prev);
            if (j < 4) gw[AA_GRAM_WORDS + j] = f2w(rg[j]);
            else rw[j - 4] = f2w(rg[j]);
        }
        rw[3] = 0u;
        H.template st<32>(AA_OFF_GRAM, gw);
        H.template st<4>(AA_OFF_RG + 4, rw);
    }
    // normal equations (FP32 Cholesky: the Gram matrix is made of FP32 sums anyway, and
    // measured iteration counts are the same as with an FP64 factorisation); columns that are
    // not valid are cut out
    float A[AA_GRAM_WORDS], r[AA_M];
    float tr = 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) {
        const bool vi = (aa.mask >> i) & 1u;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const bool vj = (aa.mask >> j) & 1u;
            A[(i * (i + 1)) / 2 + j] = (vi && vj) ? gram[(i * (i + 1)) / 2 + j] : 0.f;
        }
        r[i] = vi ? rg[i] : 0.f;
        tr += A[(i * (i + 1)) / 2 + i];
    }
    bool pd = tr > 0.f;
#pragma unroll
    for (int i = 0; i < AA_M; ++i) A[(i * (i + 1)) / 2 + i] += 1e-6f * tr + (((aa.mask >> i) & 1u) ? 0.f : 1.f);
    // A = L L' in place (L[i][i] holds the RECIPROCAL pivot), then two triangular solves.
    // All loops have constant bounds with guards, so they unroll completely and A, r
    // stay in registers.
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
    bool ok = active && aa.mask != 0u && pd;
    float fmx = 0.f;
#pragma unroll
    for (int j = 0; j < AA_M; ++j) {
        ok = ok && isfinite(r[j]);
        fmx = fmaxf(fmx, fabsf(r[j]));
    }
    // coefficients as FP16 numbers with a common power-of-two scale fs (max |gamma_j| fs in [0.5, 1))
    float ifs = 1.f;
    uint32_t fh[AA_M];
    {
        int e = (int)((f2w(fmx) >> 23) & 0xffu) - 126;
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        const float fs = w2f((uint32_t)(127 - e) << 23);
        ifs = w2f((uint32_t)(127 + e) << 23);
#pragma unroll
        for (int j = 0; j < AA_M; ++j) {
            const float f = (ok && ((aa.mask >> j) & 1u)) ? r[j] * fs : 0.f;
            fh[j] = pack_h2(f, f);
        }
    }
    H.wait_st();

    // ---- B. extrapolate optimistically: M -= sum_j gamma_j (dM_j + dG_j); remember g_k
    //         and the step; measure |step| against |g| ---------------------------------
    float ng = 0.f, ns = 0.f;
#pragma unroll 1
    for (int c = 0; c < AA_CHUNKS; ++c) {
        uint32_t cw[28], gkw[8], spw[4];
        aa_ld_cols(H, AA_OFF_DS, c, cw);
        H.wait_ld();
        float adj[8], stp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) adj[u] = 0.f;
#pragma unroll
        for (int j = 0; j < AA_M; ++j)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                adj[2 * w] = hfma_ll(cw[4 * j + w], fh[j], adj[2 * w]);
                adj[2 * w + 1] = hfma_hl(cw[4 * j + w], fh[j], adj[2 * w + 1]);
            }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            adj[u] *= ifs;
            const float gf = active ? (float)G[c * 8 + u] : 0.f;
            gkw[u] = f2w(gf);
            stp[u] = gf - adj[u];
            if (ok) M[c * 8 + u] -= (RT)adj[u];
            ng = fmaf(gf, gf, ng);
            ns = fmaf(stp[u], stp[u], ns);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) spw[w] = pack_h2(stp[2 * w] * sc, stp[2 * w + 1] * sc);
        H.template st<8>(AA_OFF_GP + 8 * c, gkw);
        H.template st<4>(AA_OFF_SP + 4 * c, spw);
    }
    H.wait_st();
    // ---- C. (rare) the extrapolated step is longer than 10 |g|: undo, keep the plain
    //         step, drop the history -----------------------------------------------------
    const bool reject = ok && !(ns <= AA_MAX_STEP2 * ng);
    if (H.any(reject)) {
#pragma unroll 1
        for (int c = 0; c < AA_CHUNKS; ++c) {
            uint32_t cw[28], gkw[8], spw[4];
            aa_ld_cols(H, AA_OFF_DS, c, cw);
            H.template ld<8>(AA_OFF_GP + 8 * c, gkw);
            H.template ld<4>(AA_OFF_SP + 4 * c, spw);
            H.wait_ld();
            float adj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) adj[u] = 0.f;
#pragma unroll
            for (int j = 0; j < AA_M; ++j)
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    adj[2 * w] = hfma_ll(cw[4 * j + w], fh[j], adj[2 * w]);
                    adj[2 * w + 1] = hfma_hl(cw[4 * j + w], fh[j], adj[2 * w + 1]);
                }
            if (reject) {
#pragma unroll
                for (int u = 0; u < 8; ++u) M[c * 8 + u] += (RT)(adj[u] * ifs);
#pragma unroll
                for (int w = 0; w < 4; ++w) spw[w] = pack_h2(w2f(gkw[2 * w]) * sc, w2f(gkw[2 * w + 1]) * sc);
            }
            H.template st<4>(AA_OFF_SP + 4 * c, spw);
        }
        H.wait_st();
    }
    if (active && aa.mask != 0u && (!ok || reject)) aa.mask = 0u;
    aa.have_prev = active;
    aa.scale_prev = sc;
}

#if defined(CVX_AA_V1)
#define aa_step aa_step_v1
#endif

// ---------------------------------------------------------------------------------
// 3x3 orthogonal factor  U Vh  of the SVD of a row-major 3x3 matrix (cvxpnpl.py:510-511,
// SVD projection WITHOUT determinant correction).  One-sided (Hestenes) Jacobi:
// rotate pairs of columns of A until they are orthogonal, A W = U S; then
// U Vh = normalise_columns(A W) W'.  Always returns an orthogonal matrix, also for
// the ill-conditioned candidates the multi-solution branch can produce (a column
// with vanishing singular value is completed by a cross product, as any SVD would
// complete it up to sign).
// ---------------------------------------------------------------------------------
CVX_HD void polar3(const double A[9], double R[9])
{
    double Bm[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int i = 0; i < 9; ++i) Bm[i] = A[i];
    for (int sweep = 0; sweep < 20; ++sweep) {
        double rot = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            const double al = Bm[p] * Bm[p] + Bm[3 + p] * Bm[3 + p] + Bm[6 + p] * Bm[6 + p];
            const double be = Bm[q] * Bm[q] + Bm[3 + q] * Bm[3 + q] + Bm[6 + q] * Bm[6 + q];
            const double ga = Bm[p] * Bm[q] + Bm[3 + p] * Bm[3 + q] + Bm[6 + p] * Bm[6 + q];
            if (!(fabs(ga) > 1e-17 * sqrt(al * be)) || !(fabs(ga) > 1e-300)) continue;
            rot += 1.0;
            const double zeta = (be - al) / (2.0 * ga);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
            const double c = 1.0 / sqrt(fma(t, t, 1.0)), sn = c * t;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double bp = Bm[3 * r + p], bq = Bm[3 * r + q];
                Bm[3 * r + p] = c * bp - sn * bq;
                Bm[3 * r + q] = sn * bp + c * bq;
                const double wp = W[3 * r + p], wq = W[3 * r + q];
                W[3 * r + p] = c * wp - sn * wq;
                W[3 * r + q] = sn * wp + c * wq;
            }
        }
        if (rot == 0.0) break;
    }
    double sv[3], smax = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        sv[j] = sqrt(Bm[j] * Bm[j] + Bm[3 + j] * Bm[3 + j] + Bm[6 + j] * Bm[6 + j]);
        smax = fmax(smax, sv[j]);
    }
    bool good[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        good[j] = sv[j] > 1e-13 * smax && sv[j] > 0.0;
        const double inv = good[j] ? 1.0 / sv[j] : 0.0;
        Bm[j] *= inv; Bm[3 + j] *= inv; Bm[6 + j] *= inv;
    }
    // complete missing left singular vectors
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (good[j]) continue;
        const int a = (j + 1) % 3, b = (j + 2) % 3;
        if (!good[a] && !good[b]) {
            // rank <= 1: pick any unit vector orthogonal to the remaining good one later
            Bm[j] = (j == 0); Bm[3 + j] = (j == 1); Bm[6 + j] = (j == 2);
            good[j] = true;
            continue;
        }
        if (!good[a] || !good[b]) {
            // one good column g: take the coordinate axis least aligned with it, orthogonalise
            const int g = good[a] ? a : b;
            const double gx = Bm[g], gy = Bm[3 + g], gz = Bm[6 + g];
            double ex = 0, ey = 0, ez = 0;
            if (fabs(gx) <= fabs(gy) && fabs(gx) <= fabs(gz)) ex = 1; else if (fabs(gy) <= fabs(gz)) ey = 1; else ez = 1;
            const double dt = ex * gx + ey * gy + ez * gz;
            double ux = ex - dt * gx, uy = ey - dt * gy, uz = ez - dt * gz;
            const double nu = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
            Bm[j] = ux * nu; Bm[3 + j] = uy * nu; Bm[6 + j] = uz * nu;
            good[j] = true;
            continue;
        }
        Bm[j] = Bm[3 + a] * Bm[6 + b] - Bm[6 + a] * Bm[3 + b];
        Bm[3 + j] = Bm[6 + a] * Bm[b] - Bm[a] * Bm[6 + b];
        Bm[6 + j] = Bm[a] * Bm[3 + b] - Bm[3 + a] * Bm[b];
        good[j] = true;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = Bm[3 * i] * W[3 * j] + Bm[3 * i + 1] * W[3 * j + 1] + Bm[3 * i + 2] * W[3 * j + 2];
}

// From a 9-vector r_c (column-major vec of a near-rotation): SO(3)/O(3) projection,
// translation t = -B r, objective r'Qr.  Writes R (row-major, world->camera) and t.
// cvxpnpl.py:510-513, 520.
template <class QIn, class BIn>
CVX_HD double finish_pose(const double rc[9], QIn Q, BIn Bm, double* R_out, double* t_out)
{
    double Rp[9];
    polar3(rc, Rp);  // Rp = projection of r_c.reshape(3,3) (row-major) = R' ; r = Rp.ravel()
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s = fma(Bm[9 * i + k], Rp[k], s);
        t_out[i] = -s;
    }
    // returned R is the transpose (cvxpnpl.py:520)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R_out[3 * i + j] = Rp[3 * j + i];
    double obj = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s = fma(Q[sidx(i, k)], Rp[k], s);
        obj = fma(s, Rp[i], obj);
    }
    return obj;
}

}  // namespace cvx
