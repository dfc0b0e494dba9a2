// tests/host/host_harness.cpp -- DEBUG/TEST ONLY.  Compiles the per-problem device
// routines (cvxpnpl_b200/csrc/*.cuh, written __host__ __device__) with the host
// compiler so the numerics of the CUDA path can be exercised by the CPU test suite
// (`-m "not gpu"`) where no GPU exists.  It is NOT part of the product, is never
// loaded by cvxpnpl_b200/, and is not a fallback: the product library fails loudly
// without CUDA.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cstdio>
#include "../../cvxpnpl_b200/csrc/pnpl_solve.cuh"
#include "../../cvxpnpl_b200/csrc/pnpl_track.cuh"
#include <cstdio>

extern "C" int host_solve(int64_t B, int n_pts, int n_lines, const double* K, int k_batched,
                          const double* pts_2d, const double* pts_3d, const double* line_2d,
                          const double* line_3d, double eps, int max_iters, int sweeps, double rho_rel,
                          double alpha, double sigma, int anderson, int variant, int fp32_iters, double* R, double* t, int32_t* n_poses, int32_t* status,
                          int32_t* iters, double* obj, double* Z)
{
    cvx::Opts o;
    o.eps2 = eps * eps;
    o.alpha = alpha;
    o.rho_rel = rho_rel;
    o.max_iters = max_iters > 0 ? max_iters : 2500;
    o.sweeps = sweeps > 0 ? sweeps : 1;
    o.sigma = sigma;
    cvx::default_params(n_pts, o.rho_rel, o.alpha, o.sigma);
    o.anderson = anderson != 0;
    o.rowk = variant == 1 ? 0.0 : 1.0;
    o.kappa = cvx::default_kappa(n_pts);
    o.early = cvx::default_early(n_pts, n_lines);
    o.aa_on2 = (anderson > 1) ? (1e-3 * anderson) * (1e-3 * anderson) : cvx::AA_RES2_ON;   // test hook: threshold in 1e-3 units
    std::vector<double> V(100), M(56), T(56), L(10), qr(45);   // T[55] = 0: zero pad for aa_step
    std::vector<uint32_t> hist(cvx::AA_WORDS, 0u);
    for (int64_t b = 0; b < B; ++b) {
        cvx::Problem pr;
        pr.K = k_batched ? K + 9 * b : K;
        pr.pts_2d = pts_2d + b * 2 * n_pts;
        pr.pts_3d = pts_3d + b * 3 * n_pts;
        pr.line_2d = line_2d + b * 4 * n_lines;
        pr.line_3d = line_3d + b * 6 * n_lines;
        pr.n_pts = n_pts;
        pr.n_lines = n_lines;
        cvx::Result rs;
        if (fp32_iters > 0) {
            // "fp32 ADMM" first phase, exactly as admm32_kernel / ortho_kernel / problem_begin_warm do it
            double pre[cvx::PRE_DOUBLES], warm[cvx::WARM_DOUBLES];
            cvx::assemble_scaled(pr, o, pre);
            cvx::start_decomposition(pre, o, cvx::Arr<1>{V.data()});
            std::vector<float> Vf(100), Mf(56), Tf(56), Lf(10), qf(45);
            cvx::ArrT<1, float> aV{Vf.data()}, aM{Mf.data()}, aT{Tf.data()}, aL{Lf.data()}, aq{qf.data()};
            cvx::problem_begin32(pre, o, aV, aM, aL, aq);
            int it = 0;
            if (std::isfinite(pre[45]))
                for (;;) {
                    const float res = cvx::pass32(o, aV, aM, aT, aL, aq);
                    ++it;
                    if (!(res > (float)cvx::FP32_EXIT_RES2) || it >= fp32_iters || it >= cvx::RESCALE_AT - 1) break;
                }
            cvx::problem_export32(aV, aM, aL, it, warm);
            cvx::warm_orthonormalise(warm);
            cvx::LaneState st;
            cvx::Arr<1> dV{V.data()}, dM{M.data()}, dT{T.data()}, dL{L.data()}, dq{qr.data()};
            cvx::problem_begin_warm(pre, warm, o, dV, dM, dL, dq, st);
            const cvx::HistMem H{hist.data(), 1};
            int wslot = 0;
            for (int guard = 0; guard < o.max_iters + 40; ++guard) {
                const bool want = cvx::pass_dr(o, dV, dM, dT, dL, dq, st);
                if (want) cvx::aa_step(dM, dT, H, st.aa, want, wslot, (float)st.res_prev);
                wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
                if (cvx::pass_eig(o, dV, dM, dT, dL, dq, st)) break;
            }
            cvx::problem_finish(pr, o, dV, dM, dT, dL, dq, st, R + b * 36, t + b * 12, Z ? Z + b * 100 : nullptr, rs);
        } else
        cvx::solve_problem(pr, o, cvx::Arr<1>{V.data()}, cvx::Arr<1>{M.data()}, cvx::Arr<1>{T.data()},
                           cvx::Arr<1>{L.data()}, cvx::Arr<1>{qr.data()}, cvx::HistMem{hist.data(), 1},
                           R + b * 36, t + b * 12, Z ? Z + b * 100 : nullptr, rs);
        n_poses[b] = rs.n_poses;
        status[b] = rs.status;
        iters[b] = rs.iters;
        if (obj) {
            obj[2 * b] = rs.pobj;
            obj[2 * b + 1] = rs.dobj;
        }
    }
    return 0;
}

extern "C" int host_extract(const double* Zin, const double* Q9, const double* Bm, double* R, double* t,
                            int32_t* status)
{
    std::vector<double> V(100), T(55), Qs(45);
    for (int i = 0; i < 10; ++i) {
        for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j);
        for (int j = 0; j <= i; ++j) T[cvx::sidx(i, j)] = 0.5 * (Zin[10 * i + j] + Zin[10 * j + i]);
    }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j <= i; ++j) Qs[cvx::sidx(i, j)] = Q9[9 * i + j];
    cvx::Arr<1> Va{V.data()}, Ta{T.data()};
    for (int s = 0; s < 40; ++s) {
        double dg = 0;
        for (int j = 0; j < 10; ++j) dg += T[cvx::sidx(j, j)] * T[cvx::sidx(j, j)];
        if (!(cvx::jacobi_sweep(Ta, Va) > 1e-32 * dg)) break;
    }
    double lam[10];
    for (int j = 0; j < 10; ++j) lam[j] = T[cvx::sidx(j, j)];
    int32_t st = 0;
    double pobj;
    int n = cvx::extract_poses(Va, lam, Qs.data(), Bm, st, 0.0, -1.0, R, t, pobj);
    *status = st;
    return n;
}

extern "C" int host_quartic(const double* c, double* x) { return cvx::quartic_real_parts(c, x); }

// plateau detector of pass_dr / the warp kernel (pnpl_solve.cuh): returns the jump length (0: none)
extern "C" int host_plateau_update(int32_t* plat, double res2, double res2_prev) { return cvx::plateau_update(*plat, res2, res2_prev); }


// The tracked solver (pnpl_track.cuh) exactly as the CUDA kernels chain it: pre-pass (assembly, start
// decomposition, Opts::early full iterations) -> tracked DR loop with Anderson steps -> park -> extraction.
// A problem whose certificate fails continues like the warp-per-problem kernel does: cold full
// decomposition of its M, full-decomposition DR loop.  fallbacks[b] = 1 for those.
extern "C" int host_solve_track(int64_t B, int n_pts, int n_lines, const double* K, int k_batched,
                                const double* pts_2d, const double* pts_3d, const double* line_2d,
                                const double* line_3d, double eps, int max_iters, double rho_rel, double alpha,
                                double sigma, int anderson, int variant, double* R, double* t, int32_t* n_poses,
                                int32_t* status, int32_t* iters, double* obj, double* Z, int32_t* fallbacks)
{
    cvx::Opts o;
    o.eps2 = eps * eps;
    o.alpha = alpha;
    o.rho_rel = rho_rel;
    o.max_iters = max_iters > 0 ? max_iters : 2500;
    o.sweeps = 1;
    o.sigma = sigma;
    cvx::default_params(n_pts, o.rho_rel, o.alpha, o.sigma);
    o.anderson = anderson != 0;
    o.rowk = variant == 1 ? 0.0 : 1.0;
    o.kappa = cvx::default_kappa(n_pts);
    o.aa_on2 = cvx::AA_RES2_ON;
    o.early = cvx::default_early(n_pts, n_lines);
    std::vector<double> V(100), M(56), T(56), L(10), qr(45), U(20), TH(2), BS(36), Qs(45), Bs(27);
    std::vector<uint32_t> hist(cvx::AA_WORDS, 0u);
    const cvx::AnyLane any;
    for (int64_t b = 0; b < B; ++b) {
        cvx::Problem pr;
        pr.K = k_batched ? K + 9 * b : K;
        pr.pts_2d = pts_2d + b * 2 * n_pts;
        pr.pts_3d = pts_3d + b * 3 * n_pts;
        pr.line_2d = line_2d + b * 4 * n_lines;
        pr.line_3d = line_3d + b * 6 * n_lines;
        pr.n_pts = n_pts;
        pr.n_lines = n_lines;
        cvx::Arr<1> aV{V.data()}, aM{M.data()}, aT{T.data()}, aL{L.data()}, aQ{qr.data()}, aU{U.data()}, aTH{TH.data()},
            aBS{BS.data()};
        double rec[cvx::PRE_DOUBLES], park[cvx::PARK_DOUBLES];
        cvx::assemble_scaled(pr, o, rec);
        cvx::start_decomposition(rec, o, aV);
        cvx::track_early(rec, o, aV, aM, aT, aL);
        cvx::LaneState st;
        cvx::track_begin(rec, o, aM, aU, aTH, aQ, st);
        const cvx::HistMem H{hist.data(), 1};
        T[55] = 0.0;
        int wslot = 0;
        int32_t it_out = 0;
        bool handed = false;
        for (int guard = 0; guard < o.max_iters + 40; ++guard) {
            const bool want = cvx::track_pass_dr(o, aM, aT, aU, aTH, aQ, st);
            if (want) cvx::aa_step(aM, aT, H, st.aa, want, wslot, (float)st.res_prev);
            wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
            const int rc = cvx::track_pass_eig(o, aM, aU, aTH, aBS, aQ, st, any);
            if (rc > 0) break;
            if (rc < 0) {
                handed = true;
#if defined(CVX_FAIL_TRACE)
                {
                    double tt[55], vv[100];
                    for (int e = 0; e < 55; ++e) tt[e] = M[e];
                    for (int i = 0; i < 100; ++i) vv[i] = (i / 10 == i % 10);
                    for (int sw = 0; sw < 12; ++sw) cvx::jacobi_sweep(cvx::Arr<1>{tt}, cvx::Arr<1>{vv});
                    double ev[10];
                    for (int j = 0; j < 10; ++j) ev[j] = tt[cvx::sidx(j, j)];
                    for (int a = 0; a < 10; ++a) for (int b2 = a + 1; b2 < 10; ++b2) if (ev[b2] > ev[a]) { double x = ev[a]; ev[a] = ev[b2]; ev[b2] = x; }
                    printf("FAIL b %ld it %d iterating %d res %.2e ev %.3e %.3e %.3e %.3e th %.3e %.3e\n", (long)b, st.it, (int)st.iterating, sqrt(st.res_prev), ev[0], ev[1], ev[2], ev[3], TH[0], TH[1]);
                }
#endif
                break;
            }
        }
        cvx::Result rs;
        if (handed) {
            // what straggler_kernel does: cold decomposition of M, then the full-decomposition loop
            cvx::track_handoff(aM, aQ, st, rec);
            for (int e = 0; e < 55; ++e) T[e] = M[e];
            for (int i = 0; i < 10; ++i)
                for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j);
            for (int s = 0; s < 12; ++s) {
                double dg = 0;
                for (int j = 0; j < 10; ++j) dg += T[cvx::sidx(j, j)] * T[cvx::sidx(j, j)];
                if (!(cvx::jacobi_sweep(aT, aV) > 1e-26 * dg)) break;
            }
            for (int j = 0; j < 10; ++j) L[j] = T[cvx::sidx(j, j)];
            T[55] = 0.0;
            cvx::aa_reset(st.aa);
            st.res_prev = 1e300;
            st.phase = st.iterating ? 0 : 1;
            for (int guard = 0; guard < o.max_iters + 40; ++guard) {
                const bool want = cvx::pass_dr(o, aV, aM, aT, aL, aQ, st);
                if (want) cvx::aa_step(aM, aT, H, st.aa, want, wslot, (float)st.res_prev);
                wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
                if (cvx::pass_eig(o, aV, aM, aT, aL, aQ, st)) break;
            }
            cvx::problem_park(aV, aL, st, park, &it_out);
        } else {
            cvx::track_park(o, aU, aTH, st, park, &it_out);
        }
        cvx::extract_parked(pr, o, park, aV, cvx::Arr<1>{Qs.data()}, cvx::Arr<1>{Bs.data()}, R + b * 36, t + b * 12,
                            Z ? Z + b * 100 : nullptr, rs);
        n_poses[b] = rs.n_poses;
        status[b] = rs.status;
        iters[b] = it_out;
        if (fallbacks) fallbacks[b] = handed ? 1 : 0;
#if defined(CVX_TRK_DEBUG)
        if (b == B - 1) printf("track_step calls %ld over %ld passes\n", cvx::g_track_steps, cvx::g_track_passes);
#endif
        if (obj) {
            obj[2 * b] = rs.pobj;
            obj[2 * b + 1] = rs.dobj;
        }
    }
    return 0;
}


// ---------------------------------------------------------------------------------------
// The role-split tracked solver (pnpl_track2.cuh) with two host threads standing in for the two CUDA threads of a
// problem, and a spinning barrier where the kernel has its named barrier.
// ---------------------------------------------------------------------------------------
#include <atomic>
#include <thread>
#include "../../cvxpnpl_b200/csrc/pnpl_track2.cuh"

namespace {
struct SpinBarrier {
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    void wait()
    {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(1 - s, std::memory_order_release);
        } else {
            while (sense.load(std::memory_order_acquire) == s) std::this_thread::yield();
        }
    }
};
struct PairSync {
    SpinBarrier* b;
    void operator()() const { b->wait(); }
};
struct PairVote {   // both threads hold the same value: the vote is the value (plus the barrier the kernel's vote implies)
    SpinBarrier* b;
    bool operator()(bool f) const { b->wait(); return f; }
};
struct Shared2 {
    std::vector<double> V, M, T, L, qr, U, TH, X, Qs, Bs;
    std::vector<float> XF;
    std::vector<uint32_t> hist;
    double rec[cvx::PRE_DOUBLES], park[cvx::PARK_DOUBLES];
    Shared2() : V(100), M(56), T(56), L(10), qr(45), U(20), TH(2), X(cvx::X2_DOUBLES), Qs(45), Bs(27), XF(cvx::X2_FLOATS), hist(cvx::AA_WORDS, 0u) {}
};
struct Args2 {
    int64_t B; int n_pts, n_lines; const double* K; int k_batched; const double *pts_2d, *pts_3d, *line_2d, *line_3d;
    cvx::Opts o; double *R, *t; int32_t *n_poses, *status, *iters; double *obj, *Z; int32_t* fallbacks;
};

template <int ROLE>
void role_thread(const Args2& a, Shared2& sh, SpinBarrier& bar)
{
    const cvx::Opts& o = a.o;
    cvx::Arr<1> aV{sh.V.data()}, aM{sh.M.data()}, aT{sh.T.data()}, aL{sh.L.data()}, aQ{sh.qr.data()}, aU{sh.U.data()},
        aTH{sh.TH.data()}, aX{sh.X.data()};
    cvx::ArrT<1, float> aXF{sh.XF.data()};
    const cvx::HistMem H{sh.hist.data(), 1};
    const PairSync sync{&bar};
    const PairVote vote{&bar};
    for (int64_t b = 0; b < a.B; ++b) {
        cvx::Problem pr;
        pr.K = a.k_batched ? a.K + 9 * b : a.K;
        pr.pts_2d = a.pts_2d + b * 2 * a.n_pts;
        pr.pts_3d = a.pts_3d + b * 3 * a.n_pts;
        pr.line_2d = a.line_2d + b * 4 * a.n_lines;
        pr.line_3d = a.line_3d + b * 6 * a.n_lines;
        pr.n_pts = a.n_pts;
        pr.n_lines = a.n_lines;
        if (ROLE == 0) {
            cvx::assemble_scaled(pr, o, sh.rec);
            cvx::start_decomposition(sh.rec, o, aV);
            cvx::track_early(sh.rec, o, aV, aM, aT, aL);
            sh.T[55] = 0.0;
        }
        bar.wait();
        cvx::LaneState st;
        cvx::t2_begin<ROLE>(sh.rec, o, aM, aU, aTH, aQ, st);
        bar.wait();
        int wslot = 0, rc = 0;
        for (int guard = 0; guard < o.max_iters + 40; ++guard) {
            rc = cvx::t2_pass(ROLE, o, true, aM, aT, aU, aTH, aQ, aX, aXF, H, st, wslot, sync, vote);
            wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
            bar.wait();
            if (rc != 0) break;
        }
        if (ROLE == 0) {
            int32_t it_out = 0;
            const bool handed = rc < 0;
            cvx::Result rs;
            if (handed) {
                cvx::track_handoff(aM, aQ, st, sh.rec);
                for (int e = 0; e < 55; ++e) sh.T[e] = sh.M[e];
                for (int i = 0; i < 10; ++i)
                    for (int j = 0; j < 10; ++j) sh.V[i * 10 + j] = (i == j);
                for (int s = 0; s < 12; ++s) {
                    double dg = 0;
                    for (int j = 0; j < 10; ++j) dg += sh.T[cvx::sidx(j, j)] * sh.T[cvx::sidx(j, j)];
                    if (!(cvx::jacobi_sweep(aT, aV) > 1e-26 * dg)) break;
                }
                for (int j = 0; j < 10; ++j) sh.L[j] = sh.T[cvx::sidx(j, j)];
                sh.T[55] = 0.0;
                std::fill(sh.hist.begin(), sh.hist.end(), 0u);
                cvx::aa_reset(st.aa);
                st.res_prev = 1e300;
                st.phase = st.iterating ? 0 : 1;
                int ws2 = 0;
                for (int guard = 0; guard < o.max_iters + 40; ++guard) {
                    const bool want = cvx::pass_dr(o, aV, aM, aT, aL, aQ, st);
                    if (want) cvx::aa_step(aM, aT, H, st.aa, want, ws2, (float)st.res_prev);
                    ws2 = (ws2 + 1 == cvx::AA_M) ? 0 : ws2 + 1;
                    if (cvx::pass_eig(o, aV, aM, aT, aL, aQ, st)) break;
                }
                cvx::problem_park(aV, aL, st, sh.park, &it_out);
                std::fill(sh.hist.begin(), sh.hist.end(), 0u);
            } else {
                cvx::track_park(o, aU, aTH, st, sh.park, &it_out);
            }
            cvx::extract_parked(pr, o, sh.park, aV, cvx::Arr<1>{sh.Qs.data()}, cvx::Arr<1>{sh.Bs.data()}, a.R + b * 36,
                                a.t + b * 12, a.Z ? a.Z + b * 100 : nullptr, rs);
            a.n_poses[b] = rs.n_poses;
            a.status[b] = rs.status;
            a.iters[b] = it_out;
            if (a.fallbacks) a.fallbacks[b] = handed ? 1 : 0;
            if (a.obj) {
                a.obj[2 * b] = rs.pobj;
                a.obj[2 * b + 1] = rs.dobj;
            }
        }
        bar.wait();
    }
}
}  // namespace

extern "C" int host_solve_track2(int64_t B, int n_pts, int n_lines, const double* K, int k_batched,
                                 const double* pts_2d, const double* pts_3d, const double* line_2d,
                                 const double* line_3d, double eps, int max_iters, double rho_rel, double alpha,
                                 double sigma, int anderson, int variant, double* R, double* t, int32_t* n_poses,
                                 int32_t* status, int32_t* iters, double* obj, double* Z, int32_t* fallbacks)
{
    Args2 a;
    a.B = B; a.n_pts = n_pts; a.n_lines = n_lines; a.K = K; a.k_batched = k_batched;
    a.pts_2d = pts_2d; a.pts_3d = pts_3d; a.line_2d = line_2d; a.line_3d = line_3d;
    a.R = R; a.t = t; a.n_poses = n_poses; a.status = status; a.iters = iters; a.obj = obj; a.Z = Z; a.fallbacks = fallbacks;
    cvx::Opts& o = a.o;
    o.eps2 = eps * eps;
    o.alpha = alpha;
    o.rho_rel = rho_rel;
    o.max_iters = max_iters > 0 ? max_iters : 2500;
    o.sweeps = 1;
    o.sigma = sigma;
    cvx::default_params(n_pts, o.rho_rel, o.alpha, o.sigma);
    o.anderson = anderson != 0;
    o.rowk = variant == 1 ? 0.0 : 1.0;
    o.kappa = cvx::default_kappa(n_pts);
    o.aa_on2 = cvx::AA_RES2_ON;
    o.early = cvx::default_early(n_pts, n_lines);
    Shared2 sh;
    SpinBarrier bar;
    std::thread t1([&] { role_thread<1>(a, sh, bar); });
    role_thread<0>(a, sh, bar);
    t1.join();
    return 0;
}
