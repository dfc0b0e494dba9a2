// pnpl_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of cvxpnpl_b200.
//
// One batched solve = pre-pass (assembly, start decomposition, difficulty bucket) ->
// counting sort of the work queue -> [FP32 first phase] -> persistent FP64 solver ->
// warp-per-problem straggler kernel -> resume (polish) -> finish (pose extraction).
//
// The persistent solver: one thread per pose problem, 128 problems per CTA, one CTA
// per SM.  The per-problem state -- eigenbasis V 100, DR iterate M 55, rotated matrix
// T 55, eigenvalues 10 doubles -- fills 221 KB of the SM's shared memory in a
// [element][thread] layout; during a Jacobi sweep T lives in registers; the Anderson
// history lives in tensor memory.  Q/rho (45 doubles per problem, read once per
// iteration) is parked in an L2-resident scratch.  Besides the correspondences (in) and
// the poses (out), only the small per-problem records that connect the kernels touch HBM.
// See DESIGN.md for the layout and the roofline discussion.
#include <cuda_runtime.h>

#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/cvxpnpl_b200.h"
#include "pnpl_core.cuh"
#include "pnpl_extract.cuh"
#include "pnpl_solve.cuh"
#include "pnpl_warp.cuh"
#include "pnpl_track.cuh"
#include "pnpl_track2.cuh"

namespace {

constexpr int NT = 128;                    // problems (threads) per CTA
constexpr int N_BUCKETS = 64;              // difficulty buckets of the work queue
constexpr int SMEM_DOUBLES = 221;          // V 100 + M 55 + T 55 (+1 zero pad) + lambda 10
constexpr size_t SMEM_BYTES = (size_t)NT * SMEM_DOUBLES * sizeof(double);

constexpr int MAX_DEVICES_SIDE = 64;
thread_local char g_err[512] = "";
thread_local int g_launches = 0;

// optional per-kernel timing of the last `solve` (desc.timing != 0): CUDA events on the
// caller's stream between the launches.  Slots: pre, admm32, ortho, fused, straggler,
// resume, finish, track, redecomp.
constexpr int N_TIMED = 9;
thread_local cudaEvent_t g_ev[N_TIMED + 1];
thread_local bool g_ev_ready = false;
thread_local int g_ev_slot[N_TIMED + 1];   // kernel slot that starts at event i, -1 = end
thread_local int g_ev_n = 0;

void mark(bool on, int slot, cudaStream_t st)
{
    if (!on || g_ev_n > N_TIMED) return;
    if (!g_ev_ready) {
        for (int i = 0; i <= N_TIMED; ++i) cudaEventCreate(&g_ev[i]);
        g_ev_ready = true;
    }
    g_ev_slot[g_ev_n] = slot;
    cudaEventRecord(g_ev[g_ev_n], st);
    ++g_ev_n;
}

// side stream + fork/join events of the concurrent service kernel: one set per host thread and device
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
thread_local SideStream g_side[MAX_DEVICES_SIDE];
SideStream* side_for_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES_SIDE) return nullptr;
    SideStream& s = g_side[dev];
    if (!s.stream) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            s = SideStream();
            return nullptr;
        }
    }
    return &s;
}

int fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

// Opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (device) context, not
// to the process, and host threads may call into the library concurrently on different devices:
// remember per device (bit d of a mask per kernel group) under a mutex.
constexpr int MAX_DEVICES = 64;
std::mutex g_attr_mutex;
bool g_attr_done[5][MAX_DEVICES];   // groups: 0 solve path, 1 null, 2 solve_sdp stage, 3 extract stage, 4 large-n assembly

template <class F>
cudaError_t opt_in_once(int group, F set_all)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    if (dev >= 0 && dev < MAX_DEVICES && g_attr_done[group][dev]) return cudaSuccess;
    e = set_all();
    if (e == cudaSuccess && dev >= 0 && dev < MAX_DEVICES) g_attr_done[group][dev] = true;
    return e;
}

using cvx::Opts;

// ---------------------------------------------------------------------------------
// Tensor memory as a per-thread scratchpad.  The CTA owns all 512 TMEM columns; warp
// w addresses lanes 32 (w % 4) .. +31, so each of the 128 threads has one TMEM lane
// = 512 private 32-bit words: the Anderson history (8 arrays x 64 words).  The
// tensor cores themselves are idle in this kernel (no dense contraction at 10x10);
// their memory is not.  SASS: LDTM / STTM.
// ---------------------------------------------------------------------------------
struct HistTmem {
    uint32_t base;   // TMEM address of column 0 in this warp's lane window
    template <int N>
    __device__ __forceinline__ void ld(int off, uint32_t* o) const;
    template <int N>
    __device__ __forceinline__ void st(int off, const uint32_t* v) const;
    __device__ __forceinline__ void wait_ld() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ void wait_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ bool any(bool f) const { return __any_sync(0xffffffffu, f) != 0; }
};

#define CVX_R4(o, k) "=r"(o[k]), "=r"(o[k + 1]), "=r"(o[k + 2]), "=r"(o[k + 3])
#define CVX_V4(v, k) "r"(v[k]), "r"(v[k + 1]), "r"(v[k + 2]), "r"(v[k + 3])
template <>
__device__ __forceinline__ void HistTmem::ld<4>(int off, uint32_t* o) const
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : CVX_R4(o, 0) : "r"(base + (uint32_t)off) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::ld<2>(int off, uint32_t* o) const
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];"
                 : "=r"(o[0]), "=r"(o[1]) : "r"(base + (uint32_t)off) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::st<2>(int off, const uint32_t* v) const
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};"
                 : : "r"(base + (uint32_t)off), "r"(v[0]), "r"(v[1]) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::st<16>(int off, const uint32_t* v) const
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 : : "r"(base + (uint32_t)off), CVX_V4(v, 0), CVX_V4(v, 4), CVX_V4(v, 8), CVX_V4(v, 12) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::ld<8>(int off, uint32_t* o) const
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : CVX_R4(o, 0), CVX_R4(o, 4) : "r"(base + (uint32_t)off) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::ld<16>(int off, uint32_t* o) const
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : CVX_R4(o, 0), CVX_R4(o, 4), CVX_R4(o, 8), CVX_R4(o, 12) : "r"(base + (uint32_t)off) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::ld<32>(int off, uint32_t* o) const
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : CVX_R4(o, 0), CVX_R4(o, 4), CVX_R4(o, 8), CVX_R4(o, 12), CVX_R4(o, 16), CVX_R4(o, 20), CVX_R4(o, 24),
                   CVX_R4(o, 28)
                 : "r"(base + (uint32_t)off) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::st<4>(int off, const uint32_t* v) const
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 : : "r"(base + (uint32_t)off), CVX_V4(v, 0) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::st<8>(int off, const uint32_t* v) const
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 : : "r"(base + (uint32_t)off), CVX_V4(v, 0), CVX_V4(v, 4) : "memory");
}
template <>
__device__ __forceinline__ void HistTmem::st<32>(int off, const uint32_t* v) const
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 : : "r"(base + (uint32_t)off), CVX_V4(v, 0), CVX_V4(v, 4), CVX_V4(v, 8), CVX_V4(v, 12), CVX_V4(v, 16),
                   CVX_V4(v, 20), CVX_V4(v, 24), CVX_V4(v, 28) : "memory");
}
#undef CVX_R4
#undef CVX_V4

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot_smem)
{
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(slot_smem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot_smem;
}

__device__ __forceinline__ void tmem_free_all(uint32_t taddr)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512) : "memory");
}

__device__ __forceinline__ const double* problem_K(const cvxpnpl_b200_desc& d, int64_t b)
{
    return d.k_batched ? d.K + 9 * b : d.K;
}

__device__ __forceinline__ cvx::Problem problem_at(const cvxpnpl_b200_desc& d, int64_t b)
{
    cvx::Problem pr;
    pr.K = problem_K(d, b);
    pr.pts_2d = d.pts_2d + b * 2 * d.n_pts;
    pr.pts_3d = d.pts_3d + b * 3 * d.n_pts;
    pr.line_2d = d.line_2d + b * 4 * d.n_lines;
    pr.line_3d = d.line_3d + b * 6 * d.n_lines;
    pr.n_pts = d.n_pts;
    pr.n_lines = d.n_lines;
    return pr;
}

// ---------------------------------------------------------------------------------
// Fused persistent kernel: assembly -> SDP -> extraction.  One CTA per SM, one
// problem per thread at a time; a finished lane immediately pulls the next problem
// index from `counter` (lane-level work stealing), so the iteration-count spread
// of the batch (median ~300, tail to 2500) costs no idle lanes.
// ---------------------------------------------------------------------------------
// control words at the head of the workspace (unsigned long long each)
enum {
    CTRL_NEXT = 0, CTRL_NSTRAG = 1, CTRL_STRAG_NEXT = 2, CTRL_RESUME_NEXT = 3, CTRL_NEXT32 = 4, CTRL_ACTIVE = 5,
    CTRL_NFAIL = 6,      // problems the tracked solver handed back (failed certificate / stragglers): length of fail_list
    CTRL_FAIL_NEXT = 7,  // queue head of the full-decomposition solver working on that list
    CTRL_TRK_NEXT = 8,   // queue head of the tracked solver
    CTRL_TRK_DONE = 9,   // CTAs of the tracked solver that have finished (the concurrent service kernel stops then)
    CTRL_SVC_NEXT = 10,  // tickets (entries of fail_list) taken by the concurrent service kernel
    CTRL_NSTRAG_START = 11,  // slab entries that exist when straggler_kernel starts (snapshot of CTRL_NSTRAG)
    CTRL_NSERVED = 12        // entries of fail_list the concurrent service kernel has served
};
// hand-backs of the tracked solver that nobody has served yet (valid once the service kernel has finished)
__device__ __forceinline__ unsigned long long unserved_handbacks(const unsigned long long* ctrl)
{
    const unsigned long long n = ctrl[CTRL_NFAIL], s = ctrl[CTRL_NSERVED];
    return n > s ? n - s : 0ULL;
}

// RESUME = false: the batch.  Lanes pull problems from ctrl[CTRL_NEXT]; once that queue
//   is empty, a lane whose problem is still in its DR loop `grace` passes later hands it
//   over to the warp-per-problem kernel (pnpl_warp.cuh) through a slab entry and leaves --
//   provided no more than `handoff_max` problems are still iterating on the whole GPU
//   (one warp of the straggler kernel each): the warp mapping buys latency with ~8x the
//   issue slots per iteration, so while many problems are alive this kernel keeps them.
// RESUME = true: second visit for the handed-over problems, whose DR loop has been
//   finished by straggler_kernel: polish the eigen-decomposition, park.
template <bool RESUME>
__global__ void __launch_bounds__(NT, 1)
solve_fused_kernel(cvxpnpl_b200_desc d, Opts o, unsigned long long* ctrl, const double* pre, double* park,
                   double* slab, const double* warm, const int32_t* order, int grace, int handoff_max, int64_t ws_stride,
                   int from_track)
{
    extern __shared__ double smem[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    const int64_t slot = (int64_t)blockIdx.x * NT + tid;
    // list modes: nothing handed over / handed back (the common case on well-posed batches) -> leave before the
    // tensor-memory allocation; a CTA beyond the list's length has nothing to do either
    if ((RESUME || from_track) && (unsigned long long)blockIdx.x * NT >= ctrl[RESUME ? CTRL_NSTRAG : CTRL_NFAIL]) return;
    if (from_track > 1 && unserved_handbacks(ctrl) <= (unsigned long long)(from_track - 1)) return;   // (from_track = 1 + direct_max)

    cvx::Arr<NT> V{smem + tid};
    cvx::Arr<NT> M{smem + (size_t)100 * NT + tid};
    cvx::Arr<NT> T{smem + (size_t)155 * NT + tid};
    cvx::Arr<NT> L{smem + (size_t)211 * NT + tid};
    cvx::GArr QR{d.workspace + slot, ws_stride};
    const uint32_t tmem_base = tmem_alloc_all(&tmem_slot);
    const HistTmem H{tmem_base + ((uint32_t)((tid >> 5) & 3) << 21)};   // lane window 32*(warp%4), lane field = bits 31:16
    T[55] = 0.0;   // zero pad read by the 8-word chunks of aa_step
    {
        // tensor memory comes up uninitialised: zero the history (the pad words must be 0)
        uint32_t zero[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) zero[u] = 0u;
        for (int c = 0; c < cvx::AA_WORDS / 32; ++c) H.st<32>(32 * c, zero);
        H.wait_st();
    }

    int64_t b = -1;
    bool exhausted = false;
    int drain = 0;   // passes since this lane saw the queue empty
    bool counted = false;   // this lane's problem is counted in ctrl[CTRL_ACTIVE]
    // from_track: the work is the list of problems the tracked solver handed back (order = that list, warm = their
    // freshly decomposed state, see redecomp_kernel); its length is only known on the device
    const unsigned long long n_work =
        RESUME ? ctrl[CTRL_NSTRAG] : (from_track ? ctrl[CTRL_NFAIL] : (unsigned long long)d.batch);
    const int q_next = RESUME ? CTRL_RESUME_NEXT : (from_track ? CTRL_FAIL_NEXT : CTRL_NEXT);
    if (from_track && grace >= 0 && n_work <= (unsigned long long)handoff_max) grace = 1;   // fits the warp kernel
    cvx::LaneState st;
    st.finite = false;
    st.iterating = false;
    st.converged = false;
    st.it = 0;
    st.phase = 0;
    st.rho = 0.0;
    st.dobj = 0.0;
    st.res_prev = 1e300;
    cvx::aa_reset(st.aa);
    int wslot = 0;   // warp-uniform history column
    for (;;) {
        while (b < 0 && !exhausted) {
            const unsigned long long nb = atomicAdd(ctrl + q_next, 1ULL);
            if (nb < n_work) {
                if (RESUME) {
                    b = cvx::problem_resume(slab + nb * cvx::HAND_DOUBLES, V, M, L, QR, st);
                } else {
                    b = (int64_t)order[nb];   // queue position -> problem (likely stragglers first)
                    if (b < 0) {              // (hand-back list: served by the concurrent service kernel already)
                        b = -1;
                        continue;
                    }
                    if (warm) {   // the FP32 first phase / the tracked solver has already advanced the problem
                        cvx::problem_begin_warm(pre + b * cvx::PRE_DOUBLES, warm + b * cvx::WARM_DOUBLES, o, V, M, L, QR,
                                                st);
                        if (from_track) {
                            const int fl = (int)pre[b * cvx::PRE_DOUBLES + cvx::TR_FLAGS];
                            if (st.finite && (fl & 1)) {   // its DR loop was over already: only polish + park
                                st.iterating = false;
                                st.converged = (fl & 2) != 0;
                                st.phase = 1;
                            }
                        }
                    } else {
                        cvx::problem_begin(pre + b * cvx::PRE_DOUBLES, o, V, M, L, QR, st);
                    }
                }
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(0xffffffffu, b < 0)) break;   // queue empty and every lane idle
        if (!RESUME && grace >= 0 && b >= 0 && st.iterating) {
            // hand-over: the queue is empty (nothing left to steal), this problem is still in
            // its DR loop `grace` passes later -> the warp-per-problem kernel finishes it
            if (drain > 0 || *(volatile unsigned long long*)(ctrl + q_next) >= n_work) {
                if (!counted) {
                    atomicAdd(ctrl + CTRL_ACTIVE, 1ULL);
                    counted = true;
                }
                ++drain;
            }
            if (drain > grace && *(volatile unsigned long long*)(ctrl + CTRL_ACTIVE) <= (unsigned long long)handoff_max) {
                const unsigned long long k = atomicAdd(ctrl + CTRL_NSTRAG, 1ULL);
                cvx::problem_handoff(V, M, L, QR, st, b, slab + k * cvx::HAND_DOUBLES);
                atomicAdd(ctrl + CTRL_ACTIVE, ~0ULL);   // -1
                counted = false;
                b = -1;
                exhausted = true;
            }
        }
        bool want = false;
        if (b >= 0) want = cvx::pass_dr(o, V, M, T, L, QR, st);
        __syncwarp();
        // warp-collective Anderson step: executed by all 32 lanes whenever one wants it
        if (H.any(want)) cvx::aa_step(M, T, H, st.aa, want, wslot, (float)st.res_prev);
        wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
        if (b >= 0) {
            if (cvx::pass_eig(o, V, M, T, L, QR, st)) {
                // park the eigen-decomposition; poses are extracted by finish_kernel
                cvx::problem_park(V, L, st, park + b * cvx::PARK_DOUBLES, d.iters + b);
                b = -1;
            }
            if (counted && (b < 0 || !st.iterating)) {   // DR loop over: no longer a hand-over candidate
                atomicAdd(ctrl + CTRL_ACTIVE, ~0ULL);
                counted = false;
            }
        }
    }
    tmem_free_all(tmem_base);
}

// ---------------------------------------------------------------------------------
// Tracked persistent solver (pnpl_track.cuh): same structure as solve_fused_kernel<false> -- one thread per problem,
// 128 problems per CTA, one CTA per SM, lane-level work stealing from a difficulty-ordered queue, Anderson history in
// tensor memory -- but the PSD projection comes from two tracked eigenpairs refined once per iteration instead of a
// full 10x10 decomposition.  Per-problem shared memory: M 55 | G 56 (the step; its first 36 words double as the
// Cholesky scratch of track_step between the Anderson step and the next DR step) | U 20 | TH 2 | Q/rho 45 = 178
// doubles.  The four warps meet at a barrier every pass (they share instruction-cache lines, see below).
// A problem whose certificate fails for good (a third positive eigenvalue that stays, or a slot that lost its
// vector), and a straggler still iterating `grace` passes after the queue ran dry, goes back into its pre-pass record
// and onto fail_list: redecomp_kernel + solve_fused_kernel<false>(from_track) finish those with the full
// decomposition.
// (Tried and dropped, profiles/README.md r2h-r2j: four more warps per CTA that serve the handed-back problems
// concurrently, warp per problem.  Their 36 KB of shared memory shrink the L1 from 45 to 9 KB, the kernel's ~100
// register spills per pass then miss, and the bulk got 0.85 ms slower -- more than the 0.5 ms tail it hides.)
// ---------------------------------------------------------------------------------
constexpr int TRK_SMEM_DOUBLES = 55 + 56 + 20 + 2 + 45;
constexpr size_t SMEM_TRK_BYTES = (size_t)NT * TRK_SMEM_DOUBLES * sizeof(double);
__global__ void __launch_bounds__(NT, 1)
solve_track_kernel(cvxpnpl_b200_desc d, Opts o, unsigned long long* ctrl, double* pre, double* park, const double* warm,
                   const int32_t* order, int32_t* fail_list, int grace, int handoff_max)
{
    extern __shared__ double smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int queue_dry;   // some lane of this CTA has seen the work queue empty (the lanes meet at a barrier every pass)
    const int tid = threadIdx.x;
    if (tid == 0) queue_dry = 0;
    cvx::Arr<NT> M{smem + tid};
    cvx::Arr<NT> G{smem + (size_t)55 * NT + tid};
    cvx::Arr<NT> U{smem + (size_t)111 * NT + tid};
    cvx::Arr<NT> TH{smem + (size_t)131 * NT + tid};
    cvx::Arr<NT> QR{smem + (size_t)133 * NT + tid};
    const uint32_t tmem_base = tmem_alloc_all(&tmem_slot);
    const HistTmem H{tmem_base + ((uint32_t)((tid >> 5) & 3) << 21)};
    {
        uint32_t zero[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) zero[u] = 0u;
        for (int c = 0; c < cvx::AA_WORDS / 32; ++c) H.st<32>(32 * c, zero);
        H.wait_st();
    }
    G[55] = 0.0;   // zero pad read by the 8-word chunks of aa_step (neither the step nor the Cholesky scratch reach it)
    const cvx::AnyLane any;
    int64_t b = -1;
    bool exhausted = false, dry_known = false;
    int drain = 0;
    bool counted = false;
    const unsigned long long n_work = (unsigned long long)d.batch;
    cvx::LaneState st;
    st.finite = false;
    st.iterating = false;
    st.converged = false;
    st.it = 0;
    st.phase = 0;
    st.bad = 0;
    st.rho = 0.0;
    st.dobj = 0.0;
    st.res_prev = 1e300;
    cvx::aa_reset(st.aa);
    int wslot = 0;
    for (;;) {
        if (b < 0 && !exhausted) {
            const unsigned long long nb = atomicAdd(ctrl + CTRL_TRK_NEXT, 1ULL);
            if (nb < n_work) {
                b = (int64_t)order[nb];
                if (warm)
                    cvx::track_begin_warm(pre + b * cvx::PRE_DOUBLES, warm + b * cvx::WARM_DOUBLES, o, M, U, TH, QR, st);
                else
                    cvx::track_begin(pre + b * cvx::PRE_DOUBLES, o, M, U, TH, QR, st);
            } else {
                exhausted = true;
                queue_dry = 1;
            }
        }
        // The four warps of the CTA start every pass together: they then walk the same (mostly straight-line,
        // ~100 KB) instruction stream and share the lines one of them brought into the SM's instruction cache.
        // ncu, free-running warps: stall_no_instruction 1.6 cycles per issued instruction (every warp streams the
        // loop body from L2: 3.8 TB/s of instruction fetch over the chip); in lock-step 0.08, kernel 5.07 -> 4.53 ms.
        if (__syncthreads_and(b < 0)) break;
        bool give_up = false;
        if (grace >= 0 && b >= 0 && st.iterating) {
            if (drain > 0 || dry_known) {
                if (!counted) {
                    atomicAdd(ctrl + CTRL_ACTIVE, 1ULL);
                    counted = true;
                }
                ++drain;
            }
            give_up = drain > grace && *(volatile unsigned long long*)(ctrl + CTRL_ACTIVE) <= (unsigned long long)handoff_max;
        }
        bool want = false;
        if (b >= 0 && !give_up) want = cvx::track_pass_dr(o, M, G, U, TH, QR, st);
        __syncwarp();
        dry_known = *(volatile int*)&queue_dry != 0;   // for the next pass (the flag is set in front of the barrier)
        __syncwarp();
        if (H.any(want)) cvx::aa_step(M, G, H, st.aa, want, wslot, (float)st.res_prev);
        wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
        if (b >= 0) {
            const int rc = give_up ? -1 : cvx::track_pass_eig(o, M, U, TH, G, QR, st, any);
            if (rc > 0) {
                cvx::track_park(o, U, TH, st, park + b * cvx::PARK_DOUBLES, d.iters + b);
                b = -1;
            } else if (rc < 0) {
                // record first, then the list entry the concurrent service kernel may be waiting for
                cvx::track_handoff(M, QR, st, pre + b * cvx::PRE_DOUBLES);
                __threadfence();
                *(volatile int32_t*)(fail_list + atomicAdd(ctrl + CTRL_NFAIL, 1ULL)) = (int32_t)b;
                __threadfence();
                b = -1;
                if (give_up) exhausted = true;
            }
            if (counted && (b < 0 || !st.iterating)) {
                atomicAdd(ctrl + CTRL_ACTIVE, ~0ULL);
                counted = false;
            }
        }
    }
    if (tid == 0) {
        __threadfence();
        atomicAdd(ctrl + CTRL_TRK_DONE, 1ULL);
    }
    tmem_free_all(tmem_base);
}

// ---------------------------------------------------------------------------------
// Tracked persistent solver, TWO threads per problem (pnpl_track2.cuh): 256 threads per CTA, still 128 problems per
// SM; thread p of warps 0-3 (role A) and thread p of warps 4-7 (role B) own problem slot p together.  Warp w and
// warp w + 4 address the same tensor-memory lanes (role A uses columns 0-255, role B 256-511 of each lane).  Between
// the phases of a pass the two warps of a pair meet at a named barrier of their own (64 threads); all eight warps
// meet ONCE per pass, at the vote at the top of the loop, which is enough to keep them on the same instruction-cache
// lines now that both roles run the same code (with CTA-wide barriers between all phases -- the r2o..r2bm form --
// every barrier waits for the slowest of the four pairs: 19 % of the stall samples).  Role A owns the queue, the
// hand-over decision, parking and hand-back; role B follows through a control word in shared memory.
// Shared memory per problem: the 178 doubles of solve_track_kernel + 5 doubles and 22 floats of exchange scratch.
// ---------------------------------------------------------------------------------
constexpr int NT2 = 2 * NT;
constexpr int TRK2_X = TRK_SMEM_DOUBLES;                          // exchange doubles
constexpr int TRK2_XF = TRK_SMEM_DOUBLES + cvx::X2_DOUBLES;       // exchange floats (22 floats = 11 doubles)
constexpr int TRK2_SMEM_DOUBLES = TRK2_XF + cvx::X2_FLOATS / 2;
constexpr size_t SMEM_TRK2_BYTES = (size_t)NT * TRK2_SMEM_DOUBLES * sizeof(double);

// CTA-wide barrier / votes from the two role-specific code paths: named barrier 1 with an explicit thread count (the
// hardware counts arrivals, whatever instruction they come from)
#ifdef CVX_CTA_BARRIERS
struct CtaSync {
    __device__ __forceinline__ void operator()() const { asm volatile("barrier.cta.sync 1, %0;" ::"n"(2 * NT) : "memory"); }
};
struct CtaVote {
    __device__ __forceinline__ bool operator()(bool f) const
    {
        uint32_t r;
        asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; barrier.cta.red.or.pred p, 1, %2, q; selp.u32 %0, 1, 0, p; }"
                     : "=r"(r) : "r"((uint32_t)f), "n"(2 * NT) : "memory");
        return r != 0;
    }
};
#endif
// The same for ONE pair of warps (w, w + 4): named barriers 2..5, 64 threads.  The eight warps then meet only once per
// pass (cta_vote_and at the top of the loop), which is what keeps them on the same instruction-cache lines.
struct PairSync {
    int id;
    __device__ __forceinline__ void operator()() const { asm volatile("barrier.cta.sync %0, 64;" ::"r"(id) : "memory"); }
};
struct PairVote {
    int id;
    __device__ __forceinline__ bool operator()(bool f) const
    {
        // both warps of a pair hold the same flags lane by lane, so the OR over one warp is the OR over the pair
        asm volatile("barrier.cta.sync %0, 64;" ::"r"(id) : "memory");
        return __any_sync(0xffffffffu, f) != 0;
    }
};
__device__ __forceinline__ bool cta_vote_and(bool f)
{
    uint32_t r;
    asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; barrier.cta.red.and.pred p, 1, %2, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(r) : "r"((uint32_t)f), "n"(2 * NT) : "memory");
    return r != 0;
}

// (one function for both roles, the role a run-time value: everything the two threads of a problem do alike -- most of a
// pass -- is then the SAME instructions for all eight warps, which walk them in lock-step out of one instruction stream)
#ifndef CVX_SLOW_IT
#define CVX_SLOW_IT 100
#endif
constexpr int SLOW_IT = CVX_SLOW_IT;
__device__ __forceinline__ void track2_loop(const int ROLE, const cvxpnpl_b200_desc& d, const Opts& o, unsigned long long* ctrl,
                                            double* pre, double* park, const double* warm, const int32_t* order,
                                            int32_t* fail_list, int grace, int handoff_max, int slow_it, double* smem,
                                            uint32_t tmem_base, int* queue_dry)
{
    const int p = threadIdx.x & (NT - 1);
    const int wq = (threadIdx.x >> 5) & 3;
    cvx::Arr<NT> M{smem + p};
    cvx::Arr<NT> G{smem + (size_t)55 * NT + p};
    cvx::Arr<NT> U{smem + (size_t)111 * NT + p};
    cvx::Arr<NT> TH{smem + (size_t)131 * NT + p};
    cvx::Arr<NT> QR{smem + (size_t)133 * NT + p};
    cvx::Arr<NT> X{smem + (size_t)TRK2_X * NT + p};
    cvx::ArrT<NT, float> XF{reinterpret_cast<float*>(smem + (size_t)TRK2_XF * NT) + p};
    const HistTmem H{tmem_base + ((uint32_t)wq << 21)};
    // barriers of a pass: per warp pair (measured: solver kernel 3.885 -> 3.65 ms against CTA-wide ones, which cost
    // 19 % of the stall samples -- every barrier waited for the slowest of four pairs); -DCVX_CTA_BARRIERS for A/B
#ifdef CVX_CTA_BARRIERS
    const CtaSync sync;
    const CtaVote vote;
#else
    const PairSync sync{2 + wq};
    const PairVote vote{2 + wq};
#endif
    {
        uint32_t zero[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) zero[u] = 0u;
        for (int c = 0; c < cvx::A2_ROLE_WORDS / 32; ++c) H.st<32>(ROLE * cvx::A2_ROLE_WORDS + 32 * c, zero);
        H.wait_st();
    }
    if (ROLE == 0) G[55] = 0.0;
    int64_t b = -1, b_next = -1;
    bool exhausted = false, counted = false, dry_known = false;
    int drain = 0;
    const unsigned long long n_work = (unsigned long long)d.batch;
    cvx::LaneState st;
    st.finite = false;
    st.iterating = false;
    st.converged = false;
    st.it = 0;
    st.phase = 0;
    st.bad = 0;
    st.rho = 0.0;
    st.dobj = 0.0;
    st.res_prev = 1e300;
    cvx::aa_reset(st.aa);
    int wslot = 0;
    // (Tried and dropped, profiles/README.md r2p: copying the next record in with 8-byte cp.async while the lane pair
    // sits out one pass -- 3.86 -> 4.12 ms: the idle pass and 125 LDGSTS per record cost more than the load latency
    // they take off the path.)
    for (;;) {
        bool fresh = false;
        if (ROLE == 0) {
            double ctl = -1.0;
#ifndef CVX_NO_PREFETCH
            // A problem whose DR loop is over only polishes its eigenpairs for one to four more passes: its lane takes the
            // NEXT problem off the queue now and prefetches that record, so the loads of t2_begin -- which hold the lane's
            // warp pair, and through the per-pass meeting the whole CTA, for an L2 / HBM round trip in nearly every pass --
            // find it in L1.
            if (b >= 0 && b_next < 0 && !exhausted && st.finite && !st.iterating) {
                const unsigned long long nb = atomicAdd(ctrl + CTRL_TRK_NEXT, 1ULL);
                if (nb < n_work) {
                    b_next = (int64_t)order[nb];
                    const char* r0 = reinterpret_cast<const char*>(pre + b_next * cvx::PRE_DOUBLES);
#pragma unroll
                    for (int l = 0; l < (cvx::TRK_DOUBLES * 8 + 127) / 128 + 1; ++l)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(r0 + 128 * l));
                    if (warm) {
                        const char* w0 = reinterpret_cast<const char*>(warm + b_next * cvx::WARM_DOUBLES);
#pragma unroll
                        for (int l = 0; l < (cvx::WARM_DOUBLES * 8 + 127) / 128 + 1; ++l)
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(w0 + 128 * l));
                    }
                } else {
                    exhausted = true;
                    *queue_dry = 1;
                }
            }
#endif
            if (b < 0) {
                if (b_next >= 0) {
                    b = b_next;
                    b_next = -1;
                    ctl = (double)b;
                    fresh = true;
                } else if (!exhausted) {
                    const unsigned long long nb = atomicAdd(ctrl + CTRL_TRK_NEXT, 1ULL);
                    if (nb < n_work) {
                        b = (int64_t)order[nb];
                        ctl = (double)b;
                        fresh = true;
                    } else {
                        exhausted = true;
                        *queue_dry = 1;
                    }
                }
            } else if (grace >= 0 && st.iterating) {
                if (drain > 0 || dry_known) {
                    if (!counted) {
                        atomicAdd(ctrl + CTRL_ACTIVE, 1ULL);
                        counted = true;
                    }
                    ++drain;
                }
                if (drain > grace && *(volatile unsigned long long*)(ctrl + CTRL_ACTIVE) <= (unsigned long long)handoff_max)
                    ctl = -3.0;
                // a slow problem moves to the concurrent service kernel (a warp of its own: ~3.6 us per iteration
                // instead of one ~12 us pass) if a service warp is waiting for work right now
                if (slow_it > 0 && st.it >= slow_it && st.it % 25 == 0 &&
                    *(volatile unsigned long long*)(ctrl + CTRL_SVC_NEXT) > *(volatile unsigned long long*)(ctrl + CTRL_NFAIL))
                    ctl = -3.0;
            }
            X[cvx::X2_CTL] = ctl;
        }
        // The one CTA-wide meeting of a pass.  Measured around it (solver kernel, 1e5 PnPL 8+4): CTA-wide barriers
        // between all phases 3.885 ms; this form 3.66; a second meeting in front of the eigenpair step 3.78; a meeting
        // every second / fourth pass 3.75 / 4.0; none at all (every pair leaves on its own vote) 4.28 -- pairs that
        // drift apart stop sharing instruction-cache lines.
        if (cta_vote_and(b < 0)) break;
        dry_known = *(volatile int*)queue_dry != 0;   // (read behind the barrier: the flag is set in front of it)
        const double ctl = X[cvx::X2_CTL];
        if (ROLE == 1 && b < 0 && ctl >= 0.0) {
            b = (int64_t)ctl;
            fresh = true;
        }
        const bool give_up = b >= 0 && ctl == -3.0;
        if (fresh) {
            const double* rec = pre + b * cvx::PRE_DOUBLES;
            if (warm) {
                if (ROLE == 0) cvx::t2_begin_warm<0>(rec, warm + b * cvx::WARM_DOUBLES, o, M, U, TH, QR, st);
                else cvx::t2_begin_warm<1>(rec, warm + b * cvx::WARM_DOUBLES, o, M, U, TH, QR, st);
            } else {
                if (ROLE == 0) cvx::t2_begin<0>(rec, o, M, U, TH, QR, st);
                else cvx::t2_begin<1>(rec, o, M, U, TH, QR, st);
            }
        }
        sync();
        int rc = cvx::t2_pass(ROLE, o, b >= 0 && !give_up, M, G, U, TH, QR, X, XF, H, st, wslot, sync, vote);
        wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
        if (b >= 0) {
            if (give_up) rc = -1;
            if (rc > 0) {
                if (ROLE == 0) cvx::track_park(o, U, TH, st, park + b * cvx::PARK_DOUBLES, d.iters + b);
                b = -1;
            } else if (rc < 0) {
                if (ROLE == 0) {
                    // record first, then the list entry the concurrent service kernel may be waiting for
                    cvx::track_handoff(M, QR, st, pre + b * cvx::PRE_DOUBLES);
                    __threadfence();
                    *(volatile int32_t*)(fail_list + atomicAdd(ctrl + CTRL_NFAIL, 1ULL)) = (int32_t)b;
                    __threadfence();
                    if (give_up && drain > 0) exhausted = true;
                }
                b = -1;
            }
            if (ROLE == 0 && counted && (b < 0 || !st.iterating)) {
                atomicAdd(ctrl + CTRL_ACTIVE, ~0ULL);
                counted = false;
            }
        }
    }
}

__global__ void __launch_bounds__(NT2, 1)
solve_track2_kernel(cvxpnpl_b200_desc d, Opts o, unsigned long long* ctrl, double* pre, double* park, const double* warm,
                    const int32_t* order, int32_t* fail_list, int grace, int handoff_max, int svc_on)
{
    extern __shared__ double smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int queue_dry;
    if (threadIdx.x == 0) queue_dry = 0;
    const uint32_t tmem_base = tmem_alloc_all(&tmem_slot);
    track2_loop(threadIdx.x < NT ? 0 : 1, d, o, ctrl, pre, park, warm, order, fail_list, grace, handoff_max, svc_on, smem,
                tmem_base, &queue_dry);
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctrl + CTRL_TRK_DONE, 1ULL);
    }
    tmem_free_all(tmem_base);
}

// Problems handed back by the tracked solver: cold eigen-decomposition of their DR iterate M (cyclic Jacobi,
// lane-parallel), exported in the WARM format the full-decomposition solver starts from.
constexpr int NT_R = 64;
constexpr size_t SMEM_R_BYTES = (size_t)NT_R * 100 * sizeof(double);
__global__ void __launch_bounds__(NT_R) redecomp_kernel(const unsigned long long* ctrl, const int32_t* fail_list,
                                                        const double* pre, double* warm, int direct_max)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    if (unserved_handbacks(ctrl) <= (unsigned long long)direct_max) return;   // few: straight to the warp-per-problem kernel
    for (unsigned long long k = (unsigned long long)blockIdx.x * NT_R + tid; k < ctrl[CTRL_NFAIL];
         k += (unsigned long long)gridDim.x * NT_R) {
    const int64_t b = fail_list[k];
    if (b < 0) continue;   // served by the concurrent service kernel
    const double* rec = pre + b * cvx::PRE_DOUBLES;
    double* w = warm + b * cvx::WARM_DOUBLES;
    cvx::Arr<NT_R> V{smem + tid};
    double t[55];
#pragma unroll
    for (int e = 0; e < 55; ++e) {
        t[e] = rec[cvx::TR_M + e];
        w[e] = t[e];
    }
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int s = 0; s < 12; ++s) {
        double dg = 0;
#pragma unroll
        for (int j = 0; j < 10; ++j) dg = fma(t[cvx::sidx(j, j)], t[cvx::sidx(j, j)], dg);
        if (!(cvx::jacobi_sweep_reg(t, V) > 1e-26 * dg)) break;
    }
#pragma unroll 4
    for (int e = 0; e < 100; ++e) w[55 + e] = V[e];
#pragma unroll
    for (int j = 0; j < 10; ++j) w[155 + j] = t[cvx::sidx(j, j)];
    w[165] = rec[cvx::TR_IT];
    }
}

// ---------------------------------------------------------------------------------
// FP32 first phase (desc.fp32_iters > 0; BASELINE.json configs[3]).  Same persistent,
// work-stealing structure as the FP64 solver, but the per-problem state is 221 floats:
// 256 problems per CTA, two warps per scheduler.  A problem leaves when its DR residual
// is below the threshold from which the FP64 solver accelerates (||X - Z||_F < 0.15),
// or at the FP32 iteration cap.  No Anderson steps, no tensor memory.
// ---------------------------------------------------------------------------------
constexpr int NT32 = 256;
constexpr size_t SMEM32_BYTES = (size_t)NT32 * SMEM_DOUBLES * sizeof(float);
__global__ void __launch_bounds__(NT32, 1)
admm32_kernel(cvxpnpl_b200_desc d, Opts o, unsigned long long* ctrl, const double* pre, double* warm, float* qr32,
              const int32_t* order, int64_t stride32, int cap32)
{
    extern __shared__ float smf[];
    const int tid = threadIdx.x;
    const int64_t slot = (int64_t)blockIdx.x * NT32 + tid;
    cvx::ArrT<NT32, float> V{smf + tid};
    cvx::ArrT<NT32, float> M{smf + (size_t)100 * NT32 + tid};
    cvx::ArrT<NT32, float> T{smf + (size_t)155 * NT32 + tid};
    cvx::ArrT<NT32, float> L{smf + (size_t)211 * NT32 + tid};
    cvx::GArrT<float> QR{qr32 + slot, stride32};
    const float thr2 = (float)cvx::FP32_EXIT_RES2;
    int64_t b = -1;
    bool exhausted = false, finite = false;
    int it = 0;
    for (;;) {
        if (b < 0 && !exhausted) {
            const unsigned long long nb = atomicAdd(ctrl + CTRL_NEXT32, 1ULL);
            if (nb < (unsigned long long)d.batch) {
                b = (int64_t)order[nb];
                const double* pb = pre + b * cvx::PRE_DOUBLES;
                cvx::problem_begin32(pb, o, V, M, L, QR);
                finite = isfinite(pb[45]);
                it = 0;
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(0xffffffffu, b < 0)) break;
        if (b >= 0) {
            bool done = true;
            if (finite) {
                const float res = cvx::pass32(o, V, M, T, L, QR);
                ++it;
                done = !(res > thr2) || it >= cap32;   // a NaN residual leaves as well
                // queue empty: nothing left to steal, so do not wait for the slow ones here --
                // any M is a valid DR state, the FP64 solver takes over where this one stops
                done = done || *(volatile unsigned long long*)(ctrl + CTRL_NEXT32) >= (unsigned long long)d.batch;
            }
            if (done) {
                cvx::problem_export32(V, M, L, it, warm + b * cvx::WARM_DOUBLES);
                b = -1;
            }
        }
    }
}

// exported FP32 eigenbases -> orthonormal in FP64 (lane-parallel, one thread per problem)
__global__ void __launch_bounds__(128) ortho_kernel(int64_t batch, double* warm)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    cvx::warm_orthonormalise(warm + b * cvx::WARM_DOUBLES);
}

// ---------------------------------------------------------------------------------
// Straggler kernel: one WARP per handed-over problem (pnpl_warp.cuh).  8 warps per
// CTA, ~9 KB of shared memory per warp.  Warps pull slab entries from a counter.
// ---------------------------------------------------------------------------------
constexpr int NT_W = 256;
constexpr size_t SMEM_W_BYTES = (NT_W / 32) * sizeof(cvx::WarpSmem);
// One handed-back problem solved by one warp: iterate M, Q/rho, rho and the iteration count from the problem's
// tracked record (read past L1: another SM may just have written it), eigen-decomposition by a warp-cooperative cold
// Jacobi, DR loop to the end; the result goes into a fresh slab entry (flag "DR loop finished") for the resume kernel.
__device__ __forceinline__ void warp_serve_record(cvx::WarpSmem& S, const Opts& o, int lane, const uint32_t pk[9],
                                                  unsigned long long* ctrl, const double* rec, int64_t b, double* slab,
                                                  int it_cap, unsigned n_track_ctas = 0, unsigned long long n_warps = 0)
{
    for (int e = lane; e < 100; e += 32) {
        const int r = e / 10, c = e - 10 * r;
        S.M[e] = __ldcg(rec + cvx::TR_M + cvx::sidx(r, c));
        S.Q[e] = (r < 9 && c < 9) ? __ldcg(rec + cvx::TR_Q + cvx::sidx(r, c)) : 0.0;
    }
    for (int e = lane; e < 56; e += 32) {
        S.gp[e] = S.sp[e] = S.gk[e] = 0.f;
#pragma unroll
        for (int j = 0; j < cvx::AA_M; ++j) S.dG[j][e] = S.dS[j][e] = 0.f;
    }
    if (lane < cvx::AA_GRAM_WORDS) S.gram[lane] = 0.f;
    if (lane < 16) S.dots[lane] = 0.f;
    int it = (int)__ldcg(rec + cvx::TR_IT);
    double rho = __ldcg(rec + cvx::TR_RHO);
    const int fl = (int)__ldcg(rec + cvx::TR_FLAGS);
    __syncwarp();
    cvx::warp_cold_decompose(S, lane, pk);
    bool converged = (fl & 2) != 0;
    bool over = (fl & 1) != 0;
    if (!over) {
        // it_cap > 0: the concurrent service kernel lends its warp for that many iterations at most, in chunks of 100;
        // a problem that needs more (a cap runner), or whose warp should leave because the bulk of the batch is over
        // and a long hand-back list is waiting for the wider kernels anyway, is left unfinished in the slab (flags 0)
        // and goes on in straggler_kernel
        const int last = (it_cap > 0 && it + it_cap < o.max_iters) ? it + it_cap : o.max_iters;
        for (;;) {
            const int stop = (it_cap > 0 && it + 100 < last) ? it + 100 : last;
            cvx::warp_dr_loop(S, o, lane, it, converged, rho, stop);
            over = converged || it >= o.max_iters;
            if (over || it >= last) break;
            int leave = 0;
            if (lane == 0)
                leave = *(volatile unsigned long long*)(ctrl + CTRL_TRK_DONE) >= (unsigned long long)n_track_ctas &&
                        *(volatile unsigned long long*)(ctrl + CTRL_NFAIL) > n_warps;
            if (__shfl_sync(0xffffffffu, leave, 0)) break;
        }
    }
    unsigned long long ks = 0;
    if (lane == 0) ks = atomicAdd(ctrl + CTRL_NSTRAG, 1ULL);
    ks = __shfl_sync(0xffffffffu, ks, 0);
    double* h = slab + ks * cvx::HAND_DOUBLES;
    for (int p = lane; p < 55; p += 32) {
        int r, c;
        cvx::unpack_idx(p, r, c);
        h[cvx::HO_M + p] = S.M[r * 10 + c];
        if (r < 9) h[cvx::HO_Q + p] = S.Q[r * 10 + c];
    }
    for (int e = lane; e < 100; e += 32) h[cvx::HO_V + e] = S.V[e];
    if (lane < 10) h[cvx::HO_L + lane] = S.L[lane];
    if (lane == 0) {
        h[cvx::HO_IT] = (double)it;
        h[cvx::HO_RHO] = rho;
        h[cvx::HO_FLAGS] = (converged ? 1.0 : 0.0) + (over ? 2.0 : 0.0);   // 2: the DR loop is over
        h[cvx::HO_B] = (double)b;
    }
    __syncwarp();
}

// DIRECT mode (fail_list != nullptr and the tracked solver handed back no more problems than there are warps here,
// `direct_max`): the entries are taken straight from the tracked solver's hand-back list (warp_serve_record) instead
// of going through redecomp_kernel and one pass of the thread solver first (~0.17 ms of latency for a handful of
// problems).  Entries the concurrent service kernel has already finished are marked (fail_list < -1) / flagged.
__global__ void __launch_bounds__(NT_W, 2) straggler_kernel(Opts o, unsigned long long* ctrl, double* slab,
                                                            const int32_t* fail_list, const double* pre, int direct_max)
{
    extern __shared__ double smem[];
    cvx::WarpSmem& S = reinterpret_cast<cvx::WarpSmem*>(smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned long long n_fail = fail_list ? ctrl[CTRL_NFAIL] : 0ULL;
    const bool direct = n_fail > 0 && unserved_handbacks(ctrl) <= (unsigned long long)direct_max;   // (served entries are skipped)
    // tickets: first the slab entries that exist when this kernel starts (hand-overs of the thread solver; problems the
    // service kernel gave back unfinished), then -- DIRECT mode -- the entries of the hand-back list nobody has served
    const unsigned long long n_slab = *(volatile unsigned long long*)(ctrl + CTRL_NSTRAG_START);
    const unsigned long long n = n_slab + (direct ? n_fail : 0ULL);
    uint32_t pk[9];
    cvx::sweep_tables(lane, pk);
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(ctrl + CTRL_STRAG_NEXT, 1ULL);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n) break;
        if (k >= n_slab) {
            const int64_t b = fail_list[k - n_slab];
            if (b >= 0) warp_serve_record(S, o, lane, pk, ctrl, pre + b * cvx::PRE_DOUBLES, b, slab, 0);
            continue;
        }
        double* h = slab + k * cvx::HAND_DOUBLES;
        if (((int)h[cvx::HO_FLAGS] & 2) != 0) continue;   // finished by the service kernel already
        // slab entry -> full-form matrices in shared memory
        for (int e = lane; e < 100; e += 32) {
            const int r = e / 10, c = e - 10 * r;
            S.M[e] = h[cvx::HO_M + cvx::sidx(r, c)];
            S.V[e] = h[cvx::HO_V + e];
            S.Q[e] = (r < 9 && c < 9) ? h[cvx::HO_Q + cvx::sidx(r, c)] : 0.0;
        }
        if (lane < 10) S.L[lane] = h[cvx::HO_L + lane];
        for (int e = lane; e < 56; e += 32) {
            S.gp[e] = S.sp[e] = S.gk[e] = 0.f;
#pragma unroll
            for (int j = 0; j < cvx::AA_M; ++j) S.dG[j][e] = S.dS[j][e] = 0.f;
        }
        if (lane < cvx::AA_GRAM_WORDS) S.gram[lane] = 0.f;
        if (lane < 16) S.dots[lane] = 0.f;
        int it = (int)h[cvx::HO_IT];
        double rho = h[cvx::HO_RHO];
        __syncwarp();
        bool converged = false;
        cvx::warp_dr_loop(S, o, lane, it, converged, rho, o.max_iters);
        for (int p = lane; p < 55; p += 32) {
            int r, c;
            cvx::unpack_idx(p, r, c);
            h[cvx::HO_M + p] = S.M[r * 10 + c];
            if (r < 9) h[cvx::HO_Q + p] = S.Q[r * 10 + c];   // Q / rho changes with a penalty rescale
        }
        for (int e = lane; e < 100; e += 32) h[cvx::HO_V + e] = S.V[e];
        if (lane < 10) h[cvx::HO_L + lane] = S.L[lane];
        if (lane == 0) {
            h[cvx::HO_IT] = (double)it;
            h[cvx::HO_RHO] = rho;
            h[cvx::HO_FLAGS] = (converged ? 1.0 : 0.0) + 2.0;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Concurrent service kernel: runs on a side stream BESIDE the persistent tracked solver, on the SMs that kernel's
// (smaller) grid leaves free, and finishes the problems the tracked solver hands back WHILE the bulk of the batch is
// still being solved -- the problems whose certificate failed, and the slow ones (more than SLOW_IT iterations) as
// long as the service has warps to spare.  Warp per problem (warp_serve_record); a warp takes a ticket, waits for
// that entry of fail_list to be published (the list is initialised to -1), solves it, marks the entry served
// (-2 - b).  When every CTA of the tracked solver has finished, the warps finish what they hold and leave; whatever
// is still unserved goes through the kernels after this one as before.  A wall-clock limit (50 ms) guards the wait: under a
// profiler that serialises kernels the service kernel runs alone, waits it out and leaves everything to those kernels.
// ---------------------------------------------------------------------------------
constexpr int NT_SVC = 512;
constexpr int N_SVC_CTAS = 2;       // SMs reserved for the service kernel
constexpr int N_SVC_CTAS_FEW = 2;   // ... for problems with fewer than 10 correspondences.  Their batches hold more slow
                                    // problems, and while a warp iteration took 6.2 us six SMs paid (1e5 PnP-8: 6.79 ms with two,
                                    // 5.81 with four, 5.56 with six).  At 3.55 us per iteration two SMs serve them as well and the
                                    // bulk keeps the other four (profiles/r2ce: PnP-8 5.04 / 5.11 / 5.14 ms with 2 / 4 / 6, 5 points +
                                    // 3 lines 7.04 / 7.13 / 7.18, 8 lines 7.3-8.1 / 7.8 / 9.7, PnL-6 15.9 with any).
constexpr int SVC_ITERS = 1 << 20;   // iterations a service warp may spend on one problem: no limit (a cap of 400 was
                                     // measured: the same or slower on every family once the warp iteration took 3.55 us)
constexpr size_t SMEM_SVC_BYTES = (NT_SVC / 32) * sizeof(cvx::WarpSmem);   // > half an SM: one CTA per SM
__global__ void __launch_bounds__(NT_SVC, 1) service_kernel(Opts o, unsigned long long* ctrl, int32_t* fail_list, const double* pre,
                                                            double* slab, unsigned n_track_ctas, unsigned long long limit_ns,
                                                            int svc_iters, long long slab_cap)
{
    extern __shared__ double smem[];
    cvx::WarpSmem& S = reinterpret_cast<cvx::WarpSmem*>(smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint32_t pk[9];
    cvx::sweep_tables(lane, pk);
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const unsigned long long n_warps = (unsigned long long)gridDim.x * (NT_SVC / 32);
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(ctrl + CTRL_SVC_NEXT, 1ULL);
        k = __shfl_sync(FULL, k, 0);
        int b = -1, leave = 0;
        for (;;) {
            if (lane == 0) {
                b = *(volatile int32_t*)(fail_list + k);
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                const bool done = *(volatile unsigned long long*)(ctrl + CTRL_TRK_DONE) >= (unsigned long long)n_track_ctas ||
                                  now - t0 > limit_ns;
                if (done) {
                    __threadfence();
                    b = *(volatile int32_t*)(fail_list + k);   // published just before the last CTA finished?
                    // the bulk is over: only a last handful of entries is still worth a warp here, anything more goes
                    // to the (74 times wider) kernels after this one
                    const unsigned long long n_pub = *(volatile unsigned long long*)(ctrl + CTRL_NFAIL);
                    if (b < 0 || n_pub > k + n_warps) leave = 1;
                }
                // (never reached by the families measured: the hand-over slab is nearly full -- leave the rest to the
                // kernels after this one, which recycle their entries)
                if ((long long)*(volatile unsigned long long*)(ctrl + CTRL_NSTRAG) >= slab_cap) leave = 1;
            }
            b = __shfl_sync(FULL, b, 0);
            leave = __shfl_sync(FULL, leave, 0);
            if (b >= 0 || leave) break;
            __nanosleep(1000);
        }
        if (leave) break;
        __threadfence();
        warp_serve_record(S, o, lane, pk, ctrl, pre + (int64_t)b * cvx::PRE_DOUBLES, b, slab, svc_iters, n_track_ctas, n_warps);
        if (lane == 0) {
            __threadfence();
            *(volatile int32_t*)(fail_list + k) = -2 - b;   // served (finished, or parked unfinished in the slab)
            atomicAdd(ctrl + CTRL_NSERVED, 1ULL);
        }
    }
}

// ---------------------------------------------------------------------------------
// Pre-pass kernel: correspondences -> Q/rho, rho and the eigen-decomposition of the DR
// start point (156 doubles per problem), one thread per problem, lane-parallel.
// ---------------------------------------------------------------------------------
constexpr int NT_P = 64;
constexpr size_t SMEM_P_BYTES = (size_t)NT_P * 100 * sizeof(double);   // V (T stays in registers)
constexpr size_t SMEM_PE_BYTES = (size_t)NT_P * SMEM_DOUBLES * sizeof(double);   // with the early iterations: V, M, T, lambda
// Difficulty-ordered work queue.  The slow problems of a batch are the ones whose optimal face
// is nearly flat: two small eigenvalues of Q (measured on seeded batches: 90 % of the 1 % slowest
// problems are among the quarter of the batch with the smallest second eigenvalue).  The
// eigenvalues come for free with the start decomposition, so the pre-pass drops every problem
// into one of 64 buckets by lambda_2(Q) / ||Q||_F (third-octave steps), a counting sort turns
// the buckets into a permutation, and the persistent kernel pulls the likely stragglers FIRST:
// they are then well advanced when the queue runs dry, and the straggler phase is shorter.
__device__ __forceinline__ int difficulty_bucket(const double* rec, const Opts& o)
{
    // eigenvalues of M0 = blkdiag(I/3 - kappa Q/rho, sigma^2): entries 0..8 belong to Q
    double m1 = -1e300, m2 = -1e300;   // the two largest of 1/3 - kappa lambda_j / rho
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const double l = rec[cvx::PRE_L + j];
        if (l > m1) {
            m2 = m1;
            m1 = l;
        } else if (l > m2) {
            m2 = l;
        }
    }
    const double lam2 = (1.0 / 3.0 - m2) / o.kappa * o.rho_rel;   // lambda_2(Q) / ||Q||_F
    if (!(lam2 > 0.0) || !isfinite(lam2)) return N_BUCKETS - 1;
    const int k = (int)floor((log2(lam2) + 20.0) * 3.0);
    return k < 0 ? 0 : (k > N_BUCKETS - 1 ? N_BUCKETS - 1 : k);
}

// EARLY: also run the first Opts::early DR iterations with the full decomposition (all lanes in the same phase) and
// leave the record in the tracked solver's format (pnpl_track.cuh).
template <bool EARLY>
__global__ void __launch_bounds__(NT_P) pre_kernel(cvxpnpl_b200_desc d, Opts o, double* pre, unsigned* bucket_count,
                                                   unsigned char* bucket_of, int64_t first, int64_t last)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int64_t b = first + (int64_t)blockIdx.x * NT_P + tid;
    if (b >= last) return;
    cvx::Arr<NT_P> V{smem + tid};
    double* out = pre + b * cvx::PRE_DOUBLES;
    cvx::assemble_scaled(problem_at(d, b), o, out);
    cvx::start_decomposition(out, o, V);   // eigen-decomposition of the start point (cold Jacobi)
    const int k = (o.kappa != 0.0) ? difficulty_bucket(out, o) : 0;
    bucket_of[b] = (unsigned char)k;
    atomicAdd(bucket_count + k, 1u);
    if (EARLY) {
        cvx::Arr<NT_P> M{smem + (size_t)100 * NT_P + tid};
        cvx::Arr<NT_P> T{smem + (size_t)155 * NT_P + tid};
        cvx::Arr<NT_P> L{smem + (size_t)211 * NT_P + tid};
        cvx::track_early(out, o, V, M, T, L);
    }
}

// The early iterations as a kernel of their own (track_early on the record pre_kernel<false> left): assembly and start
// decomposition need 100 doubles of shared memory per thread and run with eight warps per SM, the early iterations
// need 221 and run with four -- in one kernel (pre_kernel<true>) everything runs with four.
__global__ void __launch_bounds__(NT_P) early_kernel(Opts o, double* pre, int64_t first, int64_t last)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int64_t b = first + (int64_t)blockIdx.x * NT_P + tid;
    if (b >= last) return;
    cvx::Arr<NT_P> V{smem + tid};
    cvx::Arr<NT_P> M{smem + (size_t)100 * NT_P + tid};
    cvx::Arr<NT_P> T{smem + (size_t)155 * NT_P + tid};
    cvx::Arr<NT_P> L{smem + (size_t)211 * NT_P + tid};
    cvx::track_early(pre + b * cvx::PRE_DOUBLES, o, V, M, T, L);
}

// counting sort of the buckets -> queue order (likely stragglers first)
__global__ void bucket_scan_kernel(const unsigned* count, unsigned* offset)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned acc = 0;
        for (int k = 0; k < N_BUCKETS; ++k) {
            offset[k] = acc;
            acc += count[k];
        }
    }
}
__global__ void __launch_bounds__(256) bucket_scatter_kernel(int64_t batch, const unsigned char* bucket_of,
                                                             unsigned* offset, int32_t* order)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    order[atomicAdd(offset + bucket_of[b], 1u)] = (int32_t)b;
}

// ---------------------------------------------------------------------------------
// Finish kernel: parked eigen-decompositions -> poses, one thread per problem, fully
// lane-parallel (re-assembly, rank test, rank-1 / multi-solution recovery, SVD
// projection, translation, optimality flag, optional Z).
// ---------------------------------------------------------------------------------
constexpr int NT_F = 32;
constexpr size_t SMEM_F_BYTES = (size_t)NT_F * 72 * sizeof(double);   // Q 45 + B 27 (V is read where it is parked)
__global__ void __launch_bounds__(NT_F) finish_kernel(cvxpnpl_b200_desc d, Opts o, double* park)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int64_t b = (int64_t)blockIdx.x * NT_F + tid;
    if (b >= d.batch) return;
    cvx::Arr<NT_F> Qs{smem + tid};
    cvx::Arr<NT_F> Bs{smem + (size_t)45 * NT_F + tid};
    cvx::Result rs;
    // the eigenvectors stay in the parked record (L2): the rank-1 path reads ten of the hundred entries
    cvx::extract_parked(problem_at(d, b), o, park + b * cvx::PARK_DOUBLES, cvx::Arr<1>{park + b * cvx::PARK_DOUBLES}, Qs, Bs,
                        d.R + b * 36, d.t + b * 12, d.Z ? d.Z + b * 100 : nullptr, rs);
    d.n_poses[b] = rs.n_poses;
    d.status[b] = rs.status;
    if (d.obj) {
        d.obj[2 * b] = rs.pobj;
        d.obj[2 * b + 1] = rs.dobj;
    }
    if (d.record) {
        // packed [R0 | t0 | n_poses | status | iters] row: what leaves the GPU (all-gather / D2H)
        double* rec = d.record + b * CVXPNPL_RECORD;
        const double* R0 = d.R + b * 36;
        const double* t0 = d.t + b * 12;
#pragma unroll
        for (int i = 0; i < 9; ++i) rec[i] = R0[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) rec[9 + i] = t0[i];
        rec[12] = (double)rs.n_poses;
        rec[13] = (double)rs.status;
        rec[14] = (double)d.iters[b];
    }
}

// ---------------------------------------------------------------------------------
// "null" baseline (benchmarks/toolkit/methods/pnp.py:24-55): no SDP -- the smallest
// right singular vector of A (= smallest eigenvector of Q = A'A), projected onto
// O(3) by SVD, sign fixed to det +1, t = -B r.  One thread per problem.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT_F) null_kernel(cvxpnpl_b200_desc d)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int64_t b = (int64_t)blockIdx.x * NT_F + tid;
    if (b >= d.batch) return;
    cvx::Arr<NT_F> V{smem + tid};
    cvx::Arr<NT_F> Qs{smem + (size_t)100 * NT_F + tid};
    cvx::Arr<NT_F> Bs{smem + (size_t)145 * NT_F + tid};
    cvx::Arr<NT_F> T{smem + (size_t)172 * NT_F + tid};   // 55
    const cvx::Problem pr = problem_at(d, b);
    const bool finite = cvx::assemble(pr.K, pr.pts_2d, pr.pts_3d, pr.n_pts, pr.line_2d, pr.line_3d, pr.n_lines, Qs, Bs);
    double tr = 0;
    for (int i = 0; i < 9; ++i) tr += Qs[cvx::sidx(i, i)];
    for (int i = 0; i < 10; ++i) {
        for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j) ? 1.0 : 0.0;
        for (int j = 0; j <= i; ++j) T[cvx::sidx(i, j)] = (i < 9) ? Qs[cvx::sidx(i, j)] : ((j == 9) ? 1e3 * (tr + 1.0) : 0.0);
    }
    for (int s = 0; s < 40; ++s) {
        double dg = 0;
        for (int j = 0; j < 10; ++j) dg = fma(T[cvx::sidx(j, j)], T[cvx::sidx(j, j)], dg);
        if (!(cvx::jacobi_sweep(T, V) > 1e-32 * dg)) break;
    }
    int jmin = 0;
    double lmin = 1e300;
    for (int j = 0; j < 10; ++j) {
        const double l = T[cvx::sidx(j, j)];
        // the padded direction has a zero in rows 0..8 of its eigenvector: skip it
        if (fabs(V[90 + j]) < 0.5 && l < lmin) { lmin = l; jmin = j; }
    }
    double rc[9], Rm[9], tv[3];
    for (int i = 0; i < 9; ++i) rc[i] = V[i * 10 + jmin];
    cvx::finish_pose(rc, Qs, Bs, Rm, tv);
    // sign fix: R *= sign(det R) (pnp.py:53); t = -B vec(R) flips with it
    const double det = Rm[0] * (Rm[4] * Rm[8] - Rm[5] * Rm[7]) - Rm[1] * (Rm[3] * Rm[8] - Rm[5] * Rm[6]) +
                       Rm[2] * (Rm[3] * Rm[7] - Rm[4] * Rm[6]);
    const double sg = (det < 0) ? -1.0 : 1.0;
    double* Ro = d.R + b * 36;
    double* to = d.t + b * 12;
    for (int i = 0; i < 36; ++i) Ro[i] = nan("");
    for (int i = 0; i < 12; ++i) to[i] = nan("");
    for (int i = 0; i < 9; ++i) Ro[i] = finite ? sg * Rm[i] : nan("");
    for (int i = 0; i < 3; ++i) to[i] = finite ? sg * tv[i] : nan("");
    d.n_poses[b] = 1;
    d.status[b] = finite ? cvx::ST_OK : cvx::ST_NAN;
    if (d.iters) d.iters[b] = 0;
}

// ---------------------------------------------------------------------------------
// Large-n assembly (benchmarks/scalability/pnp.py:37-40 sweeps n up to 10 000): the
// only regime where the path is bandwidth bound (40 B per point streamed once).  A
// problem's correspondences are split into chunks; each CTA accumulates the 60
// Kronecker sums (P P' (x) W: 36, P' (x) W: 18, W: 6) of its chunk in registers,
// reduces them with warp shuffles + shared memory and adds them to the problem's
// accumulator with 60 atomics; a second kernel turns the sums into Q and B.
// The accumulators live in the caller's Q buffer (81 >= 60 doubles per problem).
// ---------------------------------------------------------------------------------
constexpr int NT_L = 256;
constexpr int LARGE_N = 256;          // correspondences per problem from which this path is used
constexpr int CHUNK_ELEMS = 4096;     // correspondences per CTA

__global__ void __launch_bounds__(NT_L) accumulate_kernel(cvxpnpl_b200_desc d, double* acc_out)
{
    __shared__ double red[NT_L / 32][60];
    const int64_t b = blockIdx.x;   // problems on grid.x (up to 2^31 - 1), chunks of one problem on grid.y
    const int n_total = d.n_pts + d.n_lines;
    const int lo = blockIdx.y * CHUNK_ELEMS;
    const int hi = min(lo + CHUNK_ELEMS, n_total);
    const double* K = problem_K(d, b);
    double Kl[9], Ki[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Kl[i] = K[i];
    cvx::inv3(Kl, Ki);
    cvx::Accum acc;
    cvx::accum_init(acc);
    const double* p2 = d.pts_2d + b * 2 * (int64_t)d.n_pts;
    const double* p3 = d.pts_3d + b * 3 * (int64_t)d.n_pts;
    const double* l2 = d.line_2d + b * 4 * (int64_t)d.n_lines;
    const double* l3 = d.line_3d + b * 6 * (int64_t)d.n_lines;
    // points: four per thread in flight (all loads issued before the arithmetic)
    const int hi_p = min(hi, d.n_pts);
    for (int e0 = lo + threadIdx.x; e0 < hi_p; e0 += 4 * NT_L) {
        double u[4], v[4], X[4], Y[4], Z[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = e0 + q * NT_L;
            const bool in = e < hi_p;
            const int ee = in ? e : e0;
            u[q] = p2[2 * ee]; v[q] = p2[2 * ee + 1];
            X[q] = in ? p3[3 * ee] : 0.0; Y[q] = in ? p3[3 * ee + 1] : 0.0; Z[q] = in ? p3[3 * ee + 2] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (e0 + q * NT_L >= hi_p) continue;
            double p[3], P[3] = {X[q], Y[q], Z[q]};
            cvx::bearing(Ki, u[q], v[q], p);
            const double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
            const double W[6] = {n2 - p[0] * p[0], -p[1] * p[0], n2 - p[1] * p[1],
                                 -p[2] * p[0],     -p[2] * p[1], n2 - p[2] * p[2]};
            cvx::accum_add(acc, P, W);
        }
    }
    for (int e = max(lo, d.n_pts) + threadIdx.x; e < hi; e += NT_L) {
        {
            const int i = e - d.n_pts;
            double a[3], c[3];
            cvx::bearing(Ki, l2[4 * i], l2[4 * i + 1], a);
            cvx::bearing(Ki, l2[4 * i + 2], l2[4 * i + 3], c);
            double n[3] = {a[1] * c[2] - a[2] * c[1], a[2] * c[0] - a[0] * c[2], a[0] * c[1] - a[1] * c[0]};
            const double inv = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            n[0] *= inv; n[1] *= inv; n[2] *= inv;
            const double W[6] = {n[0] * n[0], n[1] * n[0], n[1] * n[1], n[2] * n[0], n[2] * n[1], n[2] * n[2]};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double P[3] = {l3[6 * i + 3 * q], l3[6 * i + 3 * q + 1], l3[6 * i + 3 * q + 2]};
                cvx::accum_add(acc, P, W);
            }
        }
    }
    // flatten, warp-reduce, CTA-reduce, one atomic per sum
    double v[60];
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
        for (int e = 0; e < 6; ++e) v[6 * g + e] = acc.PPW[g][e];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int e = 0; e < 6; ++e) v[36 + 6 * g + e] = acc.PW[g][e];
#pragma unroll
    for (int e = 0; e < 6; ++e) v[54 + e] = acc.W[e];
#pragma unroll
    for (int k = 0; k < 60; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 60; ++k) red[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 60) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NT_L / 32; ++w) s += red[w][threadIdx.x];
        atomicAdd(acc_out + b * 81 + threadIdx.x, s);
    }
}

// ---------------------------------------------------------------------------------
// Large-n assembly, POINTS, with bulk-asynchronous staging (TMA 1-D bulk copies, cp.async.bulk + mbarrier).
// accumulate_kernel above reads its points with strided 8-byte loads (four points per thread in flight) and reaches
// a third of the HBM bandwidth.  Here one elected thread of a CTA keeps a ring of TMA_STAGES tiles in flight:
// each tile is two contiguous slabs of the problem's correspondence arrays (TMA_TILE x 16 B of pts_2d, TMA_TILE x 24 B
// of pts_3d) that the copy engine writes into shared memory while the 256 threads turn the previous tiles into
// their 60 running sums; a tile's mbarrier flips when its bytes have landed, a CTA barrier frees the stage again.
// Needs 16-byte aligned slabs: pts arrays 16-byte aligned and an even number of points per problem (the host
// checks; otherwise accumulate_kernel does the points as well).  Lines stay with accumulate_kernel.
// ---------------------------------------------------------------------------------
constexpr int TMA_TILE = 512;     // points per stage: 8 KB + 12 KB
#ifndef CVX_TMA_STAGES
#define CVX_TMA_STAGES 3
#endif
#ifndef CVX_TMA_THREADS
#define CVX_TMA_THREADS 128
#endif
#ifndef CVX_TMA_CTAS
#define CVX_TMA_CTAS 3
#endif
constexpr int TMA_STAGES = CVX_TMA_STAGES;
constexpr int NT_T = CVX_TMA_THREADS;   // several CTAs per SM: one streams while another reduces its sums
struct alignas(128) TmaStage {
    double p2[TMA_TILE * 2];
    double p3[TMA_TILE * 3];
};
struct TmaSmem {
    TmaStage st[TMA_STAGES];
    unsigned long long full[TMA_STAGES];
    double red[NT_T / 32][60];
};
constexpr size_t SMEM_T_BYTES = sizeof(TmaSmem);

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

__global__ void __launch_bounds__(NT_T, CVX_TMA_CTAS) accumulate_tma_kernel(cvxpnpl_b200_desc d, double* acc_out, int tiles_per_cta)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TmaSmem& S = *reinterpret_cast<TmaSmem*>(smem_raw);
    const int64_t b = blockIdx.x;
    const int n = d.n_pts;
    const int n_tiles = (n + TMA_TILE - 1) / TMA_TILE;
    const int t_lo = blockIdx.y * tiles_per_cta;
    const int t_hi = min(t_lo + tiles_per_cta, n_tiles);
    if (t_lo >= t_hi) return;
    const double* p2 = d.pts_2d + b * 2 * (int64_t)n;
    const double* p3 = d.pts_3d + b * 3 * (int64_t)n;
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // producer: tile t of this CTA into stage t % TMA_STAGES
    auto issue = [&](int t) {
        const int s = (t - t_lo) % TMA_STAGES;
        const int e0 = t * TMA_TILE;
        const int cnt = min(TMA_TILE, n - e0);
        mbar_expect_tx(&S.full[s], (unsigned)cnt * 40u);
        tma_bulk_load(S.st[s].p2, p2 + 2 * (int64_t)e0, (unsigned)cnt * 16u, &S.full[s]);
        tma_bulk_load(S.st[s].p3, p3 + 3 * (int64_t)e0, (unsigned)cnt * 24u, &S.full[s]);
    };
    if (tid == 0)
        for (int t = t_lo; t < min(t_lo + TMA_STAGES - 1, t_hi); ++t) issue(t);
    const double* K = problem_K(d, b);
    double Kl[9], Ki[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Kl[i] = K[i];
    cvx::inv3(Kl, Ki);
    cvx::Accum acc;
    cvx::accum_init(acc);
    for (int t = t_lo; t < t_hi; ++t) {
        const int k = t - t_lo, s = k % TMA_STAGES;
        // refill the stage the previous iteration has finished with (the CTA barrier at its end)
        if (tid == 0 && t + TMA_STAGES - 1 < t_hi) issue(t + TMA_STAGES - 1);
        mbar_wait(&S.full[s], (unsigned)(k / TMA_STAGES) & 1u);
        const int cnt = min(TMA_TILE, n - t * TMA_TILE);
        const double2* s2 = reinterpret_cast<const double2*>(S.st[s].p2);
        const double* s3 = S.st[s].p3;
#pragma unroll
        for (int q = 0; q < TMA_TILE / NT_T; ++q) {
            const int e = q * NT_T + tid;
            if (e < cnt) {
                const double2 uv = s2[e];
                double p[3], P[3] = {s3[3 * e], s3[3 * e + 1], s3[3 * e + 2]};
                cvx::bearing(Ki, uv.x, uv.y, p);
                const double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
                const double W[6] = {n2 - p[0] * p[0], -p[1] * p[0], n2 - p[1] * p[1],
                                     -p[2] * p[0],     -p[2] * p[1], n2 - p[2] * p[2]};
                cvx::accum_add(acc, P, W);
            }
        }
        __syncthreads();   // everyone is done with stage s: the next iteration's producer may overwrite it
    }
    // flatten, warp-reduce, CTA-reduce, one atomic per sum (as accumulate_kernel)
    double v[60];
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
        for (int e = 0; e < 6; ++e) v[6 * g + e] = acc.PPW[g][e];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int e = 0; e < 6; ++e) v[36 + 6 * g + e] = acc.PW[g][e];
#pragma unroll
    for (int e = 0; e < 6; ++e) v[54 + e] = acc.W[e];
#pragma unroll
    for (int k = 0; k < 60; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 60; ++k) S.red[warp][k] = v[k];
    __syncthreads();
    if (tid < 60) {
        double sm = 0;
#pragma unroll
        for (int w = 0; w < NT_T / 32; ++w) sm += S.red[w][tid];
        atomicAdd(acc_out + b * 81 + tid, sm);
    }
}

__global__ void finalize_kernel(cvxpnpl_b200_desc d, double* Q, double* Bmat)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.batch) return;
    cvx::Accum acc;
    const double* a = Q + b * 81;
    for (int g = 0; g < 6; ++g)
        for (int e = 0; e < 6; ++e) acc.PPW[g][e] = a[6 * g + e];
    for (int g = 0; g < 3; ++g)
        for (int e = 0; e < 6; ++e) acc.PW[g][e] = a[36 + 6 * g + e];
    for (int e = 0; e < 6; ++e) acc.W[e] = a[54 + e];
    double q[45], bm[27];
    cvx::reduce_accum(acc, q, bm);
    double* Qo = Q + b * 81;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) Qo[9 * i + j] = q[cvx::sidx(i, j)];
    for (int i = 0; i < 27; ++i) Bmat[b * 27 + i] = bm[i];
}

// ---------------------------------------------------------------------------------
// Stage kernels (parity testing of the individual reference functions)
// ---------------------------------------------------------------------------------
__global__ void assemble_kernel(cvxpnpl_b200_desc d, double* Q, double* Bmat)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.batch) return;
    double q[45], bm[27];
    cvx::assemble(problem_K(d, b), d.pts_2d + b * 2 * d.n_pts, d.pts_3d + b * 3 * d.n_pts, d.n_pts,
                  d.line_2d + b * 4 * d.n_lines, d.line_3d + b * 6 * d.n_lines, d.n_lines, q, bm);
    double* Qo = Q + b * 81;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) Qo[9 * i + j] = q[cvx::sidx(i, j)];
    for (int i = 0; i < 27; ++i) Bmat[b * 27 + i] = bm[i];
}

__global__ void __launch_bounds__(NT, 1)
solve_sdp_kernel(cvxpnpl_b200_desc d, Opts o, const double* Q, uint32_t* hist, int64_t ws_stride)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int64_t slot = (int64_t)blockIdx.x * NT + tid;
    cvx::Arr<NT> V{smem + tid};
    cvx::Arr<NT> M{smem + (size_t)100 * NT + tid};
    cvx::Arr<NT> T{smem + (size_t)155 * NT + tid};
    cvx::Arr<NT> L{smem + (size_t)211 * NT + tid};
    cvx::GArr qr{d.workspace + slot, ws_stride};
    const cvx::HistMem H{hist + slot, ws_stride};
    T[55] = 0.0;
    for (int e = 0; e < cvx::AA_WORDS; ++e) hist[slot + (int64_t)e * ws_stride] = 0u;
    for (int64_t b = slot; b < d.batch; b += (int64_t)gridDim.x * NT) {
        const double* Qi = Q + b * 81;
        double nq = 0;
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) nq = fma(Qi[9 * i + j], Qi[9 * i + j], nq);
        cvx::LaneState st;
        st.rho = o.rho_rel * sqrt(nq);
        st.finite = (st.rho > 0.0) && isfinite(st.rho);
        st.iterating = st.finite;
        st.converged = false;
        st.it = 0;
        st.phase = 0;
        st.dobj = 0.0;
            cvx::aa_reset(st.aa);
        st.res_prev = 1e300;
        const double ir = 1.0 / st.rho;
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j <= i; ++j) qr[cvx::sidx(i, j)] = (0.5 * (Qi[9 * i + j] + Qi[9 * j + i])) * ir;
        // start point M0 = blkdiag(I/3 - kappa Q/rho, sigma^2) and its eigen-decomposition
        // (this stage kernel is lane-parallel, so the cold Jacobi runs right here)
        for (int i = 0; i < 10; ++i) {
            for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j) ? 1.0 : 0.0;
            for (int j = 0; j <= i; ++j) {
                double m = (i == j) ? (i == 9 ? o.sigma * o.sigma : 1.0 / 3.0) : 0.0;
                if (i < 9 && st.finite) m = fma(-o.kappa, (double)qr[cvx::sidx(i, j)], m);
                M[cvx::sidx(i, j)] = m;
                T[cvx::sidx(i, j)] = m;
            }
        }
        if (st.finite)
            for (int s = 0; s < 10; ++s) {
                double dg = 0;
                for (int j = 0; j < 10; ++j) dg = fma(T[cvx::sidx(j, j)], T[cvx::sidx(j, j)], dg);
                if (!(cvx::jacobi_sweep(T, V) > 1e-26 * dg)) break;
            }
        for (int j = 0; j < 10; ++j) L[j] = T[cvx::sidx(j, j)];
        int wslot = 0;
        for (int guard = 0; guard < o.max_iters + 40; ++guard) {
            const bool want = cvx::pass_dr(o, V, M, T, L, qr, st);
            if (want) cvx::aa_step(M, T, H, st.aa, want, wslot, (float)st.res_prev);
            wslot = (wslot + 1 == cvx::AA_M) ? 0 : wslot + 1;
            if (cvx::pass_eig(o, V, M, T, L, qr, st)) break;
        }
        double lam[10];
        for (int j = 0; j < 10; ++j) lam[j] = L[j];
        int32_t status = cvx::ST_NAN;
        if (st.finite) {
            status = st.converged ? cvx::ST_OK : cvx::ST_MAX_ITERS;
            for (int j = 0; j < 10; ++j)
                if (!isfinite(lam[j])) status = cvx::ST_NAN;
        }
        const double dobj = (status != cvx::ST_NAN) ? st.dobj : nan("");
        if (d.Z) cvx::write_Z(V, lam, status == cvx::ST_NAN, d.Z + b * 100);
        if (d.status) d.status[b] = status;
        if (d.iters) d.iters[b] = st.it;
        if (d.obj) {
            d.obj[2 * b] = nan("");
            d.obj[2 * b + 1] = dobj;
        }
    }
}

// extraction stage: eigen-decomposition of the given Z (cold Jacobi), then the
// shared extraction routine.
constexpr int NT_X = 64;
__global__ void __launch_bounds__(NT_X)
extract_kernel(cvxpnpl_b200_desc d, const double* Z, const double* Q, const double* Bmat,
               const double* dobj_in, double eps)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int64_t b = (int64_t)blockIdx.x * NT_X + tid;
    if (b >= d.batch) return;
    cvx::Arr<NT_X> V{smem + tid};
    cvx::Arr<NT_X> Qs{smem + (size_t)100 * NT_X + tid};   // 45
    cvx::Arr<NT_X> Bs{smem + (size_t)145 * NT_X + tid};   // 27
    cvx::Arr<NT_X> T{smem + (size_t)172 * NT_X + tid};    // 55, later scratch
    const double* Zi = Z + b * 100;
    int32_t status = cvx::ST_OK;
    for (int i = 0; i < 10; ++i) {
        for (int j = 0; j < 10; ++j) V[i * 10 + j] = (i == j) ? 1.0 : 0.0;
        for (int j = 0; j <= i; ++j) {
            const double z = 0.5 * (Zi[10 * i + j] + Zi[10 * j + i]);
            if (!isfinite(z)) status = cvx::ST_NAN;
            T[cvx::sidx(i, j)] = z;
        }
    }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j <= i; ++j) Qs[cvx::sidx(i, j)] = Q[b * 81 + 9 * i + j];
    for (int i = 0; i < 27; ++i) Bs[i] = Bmat[b * 27 + i];
    double lam[10];
    if (status == cvx::ST_OK) {
        for (int s = 0; s < 40; ++s) {
            double dg = 0;
            for (int j = 0; j < 10; ++j) dg = fma(T[cvx::sidx(j, j)], T[cvx::sidx(j, j)], dg);
            const double off = cvx::jacobi_sweep(T, V);
            if (!(off > 1e-32 * dg)) break;
        }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) lam[j] = T[cvx::sidx(j, j)];
    double pobj = nan("");
    const bool have_d = dobj_in != nullptr;
    const double dobj = have_d ? dobj_in[b] : nan("");
    int np = cvx::extract_poses(V, lam, Qs, Bs, status, dobj, have_d ? eps : -1.0, d.R + b * 36,
                                d.t + b * 12, pobj);
    d.n_poses[b] = np;
    d.status[b] = status;
    if (d.obj) {
        d.obj[2 * b] = pobj;
        d.obj[2 * b + 1] = dobj;
    }
}

// FP64 FMA throughput probe: 8 independent chains per thread, 256 threads per CTA,
// 8 CTAs per SM.  Used by bench.py to measure the denominator of the fp64-pipe
// fraction it reports next to the (tiny) HBM fraction.
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int check_common(const cvxpnpl_b200_desc* d)
{
    if (!d) return fail(-1, "null descriptor");
    if (d->batch < 0) return fail(-2, "negative batch");
    if (d->n_pts < 0 || d->n_lines < 0) return fail(-3, "negative correspondence count");
    return 0;
}

Opts make_opts(const cvxpnpl_b200_desc* d)
{
    Opts o;
    const double eps = d->eps > 0 ? d->eps : 1e-9;
    o.eps2 = eps * eps;
    o.alpha = d->alpha;
    o.rho_rel = d->rho_rel;
    o.sigma = d->sigma;
    cvx::default_params(d->n_pts, o.rho_rel, o.alpha, o.sigma);
    o.anderson = d->anderson >= 0;
    o.aa_on2 = cvx::AA_RES2_ON;
    o.rowk = (d->variant == 1) ? 0.0 : 1.0;
    o.kappa = cvx::default_kappa(d->n_pts);
    o.max_iters = d->max_iters > 0 ? d->max_iters : 2500;
    o.sweeps = d->sweeps > 0 ? d->sweeps : 1;
    o.early = cvx::default_early(d->n_pts, d->n_lines);
    return o;
}

// header: 16 control words, then the difficulty buckets of the work queue (64 counts, 64 offsets; unsigned int)
constexpr int64_t WS_HEADER_DOUBLES = 16 + N_BUCKETS;   // 16 x 8 B + 2 x 64 x 4 B

// thread slots of the persistent grid: one CTA of NT threads per SM.  Without a
// CUDA device (CPU-side callers sizing buffers) a generous 256 SMs is assumed.
int64_t device_slots_all()
{
    int dev = 0, n_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) {
        cudaGetLastError();
        n_sm = 256;
    }
    return (int64_t)n_sm * NT;
}

// thread slots a batch actually uses: whole CTAs, never more than one per SM (small batches
// -- the scalar drop-in API solves one problem per call -- then need a small workspace)
int64_t device_slots_all();
int64_t device_slots(int64_t batch)
{
    const int64_t all = device_slots_all();
    const int64_t want = ((batch + NT - 1) / NT) * NT;
    return want < all ? (want > 0 ? want : NT) : all;
}

// workspace layout: [header 16 doubles | Q/rho 45 x slots doubles | parked results 112 x batch doubles |
//                    pre-pass 46 x batch doubles | hand-over slab 216 x slots doubles |
//                    AA history AA_WORDS x slots words (stage kernel only; the fused kernel uses TMEM) |
//                    FP32-phase export 166 x batch doubles | FP32 Q/rho 45 x 2 slots floats |
//                    queue order batch x int32 | hand-back list batch x int32 | difficulty bucket batch x byte]
size_t ws_bytes_for(int64_t slots, int64_t batch)
{
    return (WS_HEADER_DOUBLES + (size_t)slots * 45 + (size_t)batch * (cvx::PARK_DOUBLES + cvx::PRE_DOUBLES) +
            (size_t)slots * cvx::HAND_DOUBLES) *
               sizeof(double) +
           (size_t)slots * cvx::AA_WORDS * sizeof(float) +
           // FP32 first phase: exported state per problem, Q/rho scratch of its 2x wider grid
           (size_t)batch * cvx::WARM_DOUBLES * sizeof(double) + (size_t)slots * 2 * 45 * sizeof(float) +
           // queue order (int32), list of problems handed back by the tracked solver (int32) and bucket (byte) per problem
           (((size_t)batch * 9 + 15) / 16) * 16;
}

}  // namespace

extern "C" {

const char* cvxpnpl_b200_version(void) { return "cvxpnpl_b200 0.1.0 (sm_100a)"; }
const char* cvxpnpl_b200_last_error(void) { return g_err; }
int cvxpnpl_b200_last_launch_count(void) { return g_launches; }

size_t cvxpnpl_b200_workspace_bytes(int64_t batch)
{
    if (batch <= 0) return 0;
    return ws_bytes_for(device_slots(batch), batch);
}

// mode 0: the whole path; mode 1: pre-pass only, for problems [first, first + count) (the workspace
// header is reset when first == 0); mode 2 (desc.skip_prepass): everything after the pre-pass
static int solve_impl(const cvxpnpl_b200_desc* d, void* stream, int mode, int64_t first, int64_t count)
{
    g_launches = 0;
    if (int rc = check_common(d)) return rc;
    if (d->batch == 0) return 0;
    if (d->n_pts + d->n_lines <= 0) return fail(-4, "no correspondences");
    if (d->batch > 2000000000LL) return fail(-8, "at most 2e9 problems per call (32-bit queue order)");
    if (!d->K || (d->n_pts && (!d->pts_2d || !d->pts_3d)) || (d->n_lines && (!d->line_2d || !d->line_3d)))
        return fail(-5, "null input pointer");
    if (!d->R || !d->t || !d->n_poses || !d->status || !d->iters) return fail(-6, "null output pointer");
    if (d->workspace_bytes < cvxpnpl_b200_workspace_bytes(d->batch) ||
        (cvxpnpl_b200_workspace_bytes(d->batch) && !d->workspace))
        return fail(-7, "workspace too small");
    {
        const cudaError_t e = opt_in_once(0, [] {
            cudaError_t e = cudaFuncSetAttribute(solve_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)SMEM_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(solve_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(straggler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_W_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(pre_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_P_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(pre_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PE_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(early_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PE_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(solve_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_TRK_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(solve_track2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_TRK2_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(service_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SVC_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(redecomp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_R_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(admm32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM32_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_F_BYTES);
            return e;
        });
        if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    }
    const int64_t slots = device_slots(d->batch);
    // persistent grid: one CTA per SM (shared memory allows exactly one), never more
    // CTAs than there is work for
    const int64_t want = (d->batch + NT - 1) / NT;
    const int64_t blocks = want < slots / NT ? want : slots / NT;
    // control words (work counters) in front of the Q/rho scratch
    unsigned long long* ctrl = (unsigned long long*)d->workspace;
    cvxpnpl_b200_desc dd = *d;
    dd.workspace = d->workspace + WS_HEADER_DOUBLES;
    double* park = dd.workspace + slots * 45;
    double* pre = park + d->batch * cvx::PARK_DOUBLES;
    double* slab = pre + d->batch * cvx::PRE_DOUBLES;
    // hand-over grace (passes after the queue ran dry); desc.handoff: 0 default, < 0 never
    // (a batch that fits the straggler grid -- one warp per problem -- goes there at once: for
    // the scalar API that is 0.2 ms instead of 1.1 ms per call)
    const int64_t n_sm = device_slots_all() / NT;
    const int grace = d->handoff < 0 ? -1 : (d->handoff > 0 ? d->handoff : (d->batch <= n_sm * 2 * (NT_W / 32) ? 1 : 40));
    const Opts o = make_opts(d);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1 && (first < 0 || count < 0 || first + count > d->batch)) return fail(-2, "pre-pass range outside the batch");
    if (mode != 2 && (mode == 0 || first == 0)) {
        cudaError_t e0 = cudaMemsetAsync(ctrl, 0, WS_HEADER_DOUBLES * sizeof(double), st);
        if (e0 != cudaSuccess) return fail((int)e0, cudaGetErrorString(e0));
    }
    // PSD projection: tracked eigenpairs (default) or the full decomposition every iteration (desc.psd_mode = 1;
    // also for batches that fit the warp-per-problem grid, which go there at once)
    const bool tracked = d->psd_mode != 1 && d->batch > n_sm * 2 * (NT_W / 32);
    const bool early = tracked && d->fp32_iters <= 0;   // the FP32 first phase starts from the start decomposition
    const bool tm = d->timing != 0 && mode != 1;
    g_ev_n = 0;
    // regions behind the hand-over slab (see ws_bytes_for)
    double* warm = (double*)((uint32_t*)(slab + slots * cvx::HAND_DOUBLES) + slots * cvx::AA_WORDS);
    float* qr32 = (float*)(warm + d->batch * cvx::WARM_DOUBLES);
    int32_t* order = (int32_t*)(qr32 + slots * 2 * 45);
    int32_t* fail_list = order + d->batch;
    unsigned char* bucket_of = (unsigned char*)(fail_list + d->batch);
    unsigned* bucket_count = (unsigned*)(ctrl + 16);
    unsigned* bucket_offset = bucket_count + N_BUCKETS;
    mark(tm, 0, st);
    int n_pre_launches = 1;
    if (mode != 2) {
        const int64_t lo = mode == 1 ? first : 0, hi = mode == 1 ? first + count : d->batch;
        if (hi > lo) {
            static const bool fused_pre = getenv("CVXPNPL_B200_FUSED_PRE") != nullptr;   // (A/B: one kernel)
            const unsigned nb = (unsigned)((hi - lo + NT_P - 1) / NT_P);
            if (early && fused_pre) {
                pre_kernel<true><<<nb, NT_P, SMEM_PE_BYTES, st>>>(dd, o, pre, bucket_count, bucket_of, lo, hi);
            } else {
                pre_kernel<false><<<nb, NT_P, SMEM_P_BYTES, st>>>(dd, o, pre, bucket_count, bucket_of, lo, hi);
                if (early) {
                    early_kernel<<<nb, NT_P, SMEM_PE_BYTES, st>>>(o, pre, lo, hi);
                    ++n_pre_launches;
                }
            }
        }
        if (mode == 1) {
            g_launches = n_pre_launches;
            cudaError_t e1 = cudaGetLastError();
            if (e1 != cudaSuccess) return fail((int)e1, cudaGetErrorString(e1));
            return 0;
        }
    }
    bucket_scan_kernel<<<1, 32, 0, st>>>(bucket_count, bucket_offset);
    bucket_scatter_kernel<<<(unsigned)((d->batch + 255) / 256), 256, 0, st>>>(d->batch, bucket_of, bucket_offset, order);
    g_launches = (mode == 2) ? 4 : 4 + n_pre_launches;   // pre-pass, two sort kernels, solver, finish
    const double* warm_in = nullptr;
    if (d->fp32_iters > 0) {
        // FP32 first phase, then its bases made orthonormal in FP64
        const int64_t want32 = (d->batch + NT32 - 1) / NT32;
        const int64_t blocks32 = want32 < slots / NT ? want32 : slots / NT;
        mark(tm, 1, st);
        // the FP32 phase stops before the penalty rescale of slow problems (RESCALE_AT) is due
        const int cap32 = d->fp32_iters < cvx::RESCALE_AT ? d->fp32_iters : cvx::RESCALE_AT - 1;
        admm32_kernel<<<(unsigned)blocks32, NT32, SMEM32_BYTES, st>>>(dd, o, ctrl, pre, warm, qr32, order,
                                                                      blocks32 * NT32, cap32);
        mark(tm, 2, st);
        ortho_kernel<<<(unsigned)((d->batch + 127) / 128), 128, 0, st>>>(d->batch, warm);
        warm_in = warm;
        g_launches += 2;
    }
    // straggler grid: one warp per handed-over problem, at most 2 CTAs of 8 warps per SM
    const int64_t warps_per_cta = NT_W / 32;
    int64_t wblocks = (blocks * NT + warps_per_cta - 1) / warps_per_cta;   // at most one hand-over per lane
    const int64_t wcap = n_sm * 2;
    if (wblocks > wcap) wblocks = wcap;
    // (two or three problems per warp were measured too: no difference)
    const int handoff_max = (int)(wblocks * warps_per_cta);
    // a hand-back list no longer than the warp-per-problem grid goes straight there (straggler_kernel, DIRECT mode)
    const int direct_max = (tracked && grace >= 0) ? handoff_max : 0;
    if (tracked) {
        mark(tm, 7, st);
        // Stragglers of the tracked solver: one of its passes takes ~10 us, an iteration of the warp-per-problem kernel
        // ~3.6 us, and the detour (cold decomposition, hand-over, resume) costs ~0.3 ms: handing over earlier than 64
        // passes after the queue ran dry does not pay (1e5 PnPL 8+4, re-swept with the pair-barrier kernel,
        // profiles/r2cc_grace_sweep.txt: grace 24..48 5.08 ms with ~2000 hand-overs, 56 4.92, 64..100 4.81 with ~55).
        const int track_grace = d->handoff != 0 ? grace : 64;
        // The concurrent service kernel (side stream): two SMs -- left free by a grid of n_sm - 2 CTAs when the batch
        // fills the GPU -- finish handed-back problems warp per problem while the bulk is still being solved.
        // desc.handoff < 0 (no warp-per-problem kernels at all) switches it off.
        static const bool no_service = getenv("CVXPNPL_B200_NO_SERVICE") != nullptr;   // (diagnostics)
        static const int svc_iters = getenv("CVXPNPL_B200_SVC_ITERS") ? atoi(getenv("CVXPNPL_B200_SVC_ITERS")) : SVC_ITERS;
        static const int slow_it_env = getenv("CVXPNPL_B200_SLOW_IT") ? atoi(getenv("CVXPNPL_B200_SLOW_IT")) : 0;
        // a slow problem moves to an idle service warp from this iteration on: 100 where that is rare (PnPL 8+4: p99 = 74),
        // 200 for the families with fewer than 10 correspondences, whose idle warps would otherwise go to the many problems
        // that need 100-150 iterations instead of the few that need 1000 (1e5 problems, 5 points + 3 lines: 8.0 -> 7.2 ms)
        const int slow_it = slow_it_env > 0 ? slow_it_env : (d->n_pts + d->n_lines < 10 ? 2 * SLOW_IT : SLOW_IT);
        SideStream* side = (grace >= 0 && d->psd_mode != 2 && !no_service) ? side_for_current_device() : nullptr;
        const bool svc_on = side != nullptr;
        // families whose batches hold more slow problems (fewer than 10 correspondences) get a wider service
        static const int svc_ctas_env = getenv("CVXPNPL_B200_SVC_CTAS") ? atoi(getenv("CVXPNPL_B200_SVC_CTAS")) : 0;
        const int n_svc = svc_ctas_env > 0 ? svc_ctas_env : (d->n_pts + d->n_lines < 10 ? N_SVC_CTAS_FEW : N_SVC_CTAS);
        int64_t blocks_trk = blocks;
        if (svc_on && blocks_trk > n_sm - n_svc) blocks_trk = n_sm - n_svc;
        cudaMemsetAsync(fail_list, 0xFF, (size_t)d->batch * sizeof(int32_t), st);   // -1: entry not published yet
        if (svc_on) {
            cudaEventRecord(side->fork, st);
            cudaStreamWaitEvent(side->stream, side->fork, 0);
            // slab entries the service may fill: what the kernels after it can still add stays free
            const long long slab_cap = (long long)slots - 2LL * handoff_max - (long long)n_svc * (NT_SVC / 32);
            service_kernel<<<n_svc, NT_SVC, SMEM_SVC_BYTES, side->stream>>>(o, ctrl, fail_list, pre, slab, (unsigned)blocks_trk,
                                                                            50000000ULL /* 50 ms */, svc_iters, slab_cap);
            ++g_launches;
        }
        if (d->psd_mode == 2)   // one thread per problem (the round-2a form, kept for A/B runs)
            solve_track_kernel<<<(unsigned)blocks_trk, NT, SMEM_TRK_BYTES, st>>>(dd, o, ctrl, pre, park, warm_in, order,
                                                                                 fail_list, track_grace, handoff_max);
        else
            solve_track2_kernel<<<(unsigned)blocks_trk, NT2, SMEM_TRK2_BYTES, st>>>(dd, o, ctrl, pre, park, warm_in, order,
                                                                                    fail_list, track_grace, handoff_max,
                                                                                    svc_on ? slow_it : 0);
        if (svc_on) {
            cudaEventRecord(side->join, side->stream);
            cudaStreamWaitEvent(st, side->join, 0);
        }
        mark(tm, 8, st);
        {
            const int64_t want_r = (d->batch + NT_R - 1) / NT_R, cap_r = n_sm * 4;   // grid-stride over the list
            redecomp_kernel<<<(unsigned)(want_r < cap_r ? want_r : cap_r), NT_R, SMEM_R_BYTES, st>>>(ctrl, fail_list, pre, warm,
                                                                                                 direct_max);
        }
        mark(tm, 3, st);
        solve_fused_kernel<false><<<(unsigned)blocks, NT, SMEM_BYTES, st>>>(dd, o, ctrl, pre, park, slab, warm, fail_list,
                                                                            grace, handoff_max, slots, 1 + direct_max);
        g_launches += 2;
    } else {
        mark(tm, 3, st);
        solve_fused_kernel<false><<<(unsigned)blocks, NT, SMEM_BYTES, st>>>(dd, o, ctrl, pre, park, slab, warm_in, order,
                                                                            grace, handoff_max, slots, 0);
    }
    if (grace >= 0) {
        mark(tm, 4, st);
        // snapshot of the slab's length: straggler_kernel allocates further entries while it runs
        cudaMemcpyAsync(ctrl + CTRL_NSTRAG_START, ctrl + CTRL_NSTRAG, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st);
        straggler_kernel<<<(unsigned)wblocks, NT_W, SMEM_W_BYTES, st>>>(o, ctrl, slab, tracked ? fail_list : nullptr, pre,
                                                                        direct_max);
        mark(tm, 5, st);
        solve_fused_kernel<true><<<(unsigned)blocks, NT, SMEM_BYTES, st>>>(dd, o, ctrl, pre, park, slab, nullptr, nullptr,
                                                                          -1, 0, slots, 0);
        g_launches += 2;
    }
    mark(tm, 6, st);
    finish_kernel<<<(unsigned)((d->batch + NT_F - 1) / NT_F), NT_F, SMEM_F_BYTES, st>>>(dd, o, park);
    mark(tm, -1, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    return 0;
}

int cvxpnpl_b200_solve(const cvxpnpl_b200_desc* d, void* stream)
{
    return solve_impl(d, stream, (d && d->skip_prepass) ? 2 : 0, 0, 0);
}

int cvxpnpl_b200_prepass(const cvxpnpl_b200_desc* d, int64_t first, int64_t count, void* stream)
{
    return solve_impl(d, stream, 1, first, count);
}

int cvxpnpl_b200_kernel_times(float* ms, int n)
{
    if (!ms || n < N_TIMED) return fail(-5, "kernel_times needs room for 9 floats");
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    if (g_ev_n < 2) return fail(-9, "the last solve on this thread was not timed (desc.timing = 0)");
    cudaError_t e = cudaEventSynchronize(g_ev[g_ev_n - 1]);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    for (int i = 0; i + 1 < g_ev_n; ++i) {
        float t = 0.f;
        e = cudaEventElapsedTime(&t, g_ev[i], g_ev[i + 1]);
        if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
        if (g_ev_slot[i] >= 0) ms[g_ev_slot[i]] = t;
    }
    return 0;
}

int cvxpnpl_b200_null(const cvxpnpl_b200_desc* d, void* stream)
{
    g_launches = 0;
    if (int rc = check_common(d)) return rc;
    if (d->batch == 0) return 0;
    if (d->n_pts + d->n_lines <= 0) return fail(-4, "no correspondences");
    if (!d->K || (d->n_pts && (!d->pts_2d || !d->pts_3d)) || (d->n_lines && (!d->line_2d || !d->line_3d)))
        return fail(-5, "null input pointer");
    if (!d->R || !d->t || !d->n_poses || !d->status) return fail(-6, "null output pointer");
    const size_t smem = (size_t)NT_F * 227 * sizeof(double);
    {
        const cudaError_t e = opt_in_once(1, [smem] {
            return cudaFuncSetAttribute(null_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        });
        if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    }
    null_kernel<<<(unsigned)((d->batch + NT_F - 1) / NT_F), NT_F, smem, (cudaStream_t)stream>>>(*d);
    g_launches = 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    return 0;
}

int cvxpnpl_b200_fp64_probe(double* out, int64_t out_len, int iters, int64_t* flops, void* stream)
{
    g_launches = 0;
    int dev = 0, n_sm = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e0 != cudaSuccess) return fail((int)e0, cudaGetErrorString(e0));
    const int blocks = n_sm * 8, threads = 256;
    if (!out || out_len < (int64_t)blocks * threads) return fail(-7, "probe output too small");
    fp64_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999999, 1e-9);
    g_launches = 1;
    if (flops) *flops = (int64_t)blocks * threads * (int64_t)iters * 64 * 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    return 0;
}

int cvxpnpl_b200_assemble(const cvxpnpl_b200_desc* d, double* Q, double* Bmat, void* stream)
{
    g_launches = 0;
    if (int rc = check_common(d)) return rc;
    if (d->batch == 0) return 0;
    if (!Q || !Bmat || !d->K) return fail(-5, "null pointer");
    if (d->n_pts + d->n_lines <= 0) return fail(-4, "no correspondences");
    if (d->n_pts + d->n_lines >= LARGE_N) {
        // bandwidth-bound regime: chunked streaming reduction (3 operations on the stream)
        const int chunks = (d->n_pts + d->n_lines + CHUNK_ELEMS - 1) / CHUNK_ELEMS;
        if (d->batch > 2147483647LL || chunks > 65535)
            return fail(-8, "large-n assembly: at most 2^31 - 1 problems of at most 65535 x 4096 correspondences per call");
        cudaError_t e0 = cudaMemsetAsync(Q, 0, (size_t)d->batch * 81 * sizeof(double), (cudaStream_t)stream);
        if (e0 != cudaSuccess) return fail((int)e0, cudaGetErrorString(e0));
        g_launches = 0;
        // points through the TMA-staged kernel when their slabs are 16-byte aligned (desc.psd_mode = 1, "the round-1
        // path", keeps the plain-load kernel: A/B measurements)
        const bool tma = d->n_pts >= LARGE_N && (d->n_pts % 2) == 0 && ((uintptr_t)d->pts_2d % 16) == 0 &&
                         ((uintptr_t)d->pts_3d % 16) == 0 && d->psd_mode != 1;
        cvxpnpl_b200_desc dl = *d;
        if (tma) {
            e0 = opt_in_once(4, [] {
                return cudaFuncSetAttribute(accumulate_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_T_BYTES);
            });
            if (e0 != cudaSuccess) return fail((int)e0, cudaGetErrorString(e0));
            // enough CTAs to fill the GPU a few times over, at least four tiles each (the ring's depth)
            const int n_tiles = (d->n_pts + TMA_TILE - 1) / TMA_TILE;
            const int64_t n_sm = device_slots_all() / NT;
            int splits = (int)((8 * n_sm + d->batch - 1) / d->batch);
            if (splits > (n_tiles + 3) / 4) splits = (n_tiles + 3) / 4;
            if (splits < 1) splits = 1;
            const int tiles_per_cta = (n_tiles + splits - 1) / splits;
            splits = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
            accumulate_tma_kernel<<<dim3((unsigned)d->batch, (unsigned)splits), NT_T, SMEM_T_BYTES, (cudaStream_t)stream>>>(
                *d, Q, tiles_per_cta);
            ++g_launches;
            dl.n_pts = 0;   // the plain-load kernel below only has the lines left
        }
        if (dl.n_pts + dl.n_lines > 0) {
            const int chunks_l = (dl.n_pts + dl.n_lines + CHUNK_ELEMS - 1) / CHUNK_ELEMS;
            accumulate_kernel<<<dim3((unsigned)d->batch, (unsigned)chunks_l), NT_L, 0, (cudaStream_t)stream>>>(dl, Q);
            ++g_launches;
        }
        finalize_kernel<<<(unsigned)((d->batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*d, Q, Bmat);
        ++g_launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
        return 0;
    }
    const int64_t blocks = (d->batch + 127) / 128;
    assemble_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*d, Q, Bmat);
    g_launches = 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    return 0;
}

int cvxpnpl_b200_solve_sdp(const cvxpnpl_b200_desc* d, const double* Q, void* stream)
{
    g_launches = 0;
    if (int rc = check_common(d)) return rc;
    if (d->batch == 0) return 0;
    if (!Q) return fail(-5, "null Q");
    if (d->workspace_bytes < cvxpnpl_b200_workspace_bytes(d->batch) ||
        (cvxpnpl_b200_workspace_bytes(d->batch) && !d->workspace))
        return fail(-7, "workspace too small");
    {
        const cudaError_t e = opt_in_once(2, [] {
            return cudaFuncSetAttribute(solve_sdp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        });
        if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    }
    const int64_t slots = device_slots(d->batch);
    const int64_t want = (d->batch + NT - 1) / NT;
    const int64_t blocks = want < slots / NT ? want : slots / NT;
    cvxpnpl_b200_desc dd = *d;
    dd.workspace = d->workspace + WS_HEADER_DOUBLES;
    uint32_t* hist = (uint32_t*)(dd.workspace + slots * 45 + d->batch * (cvx::PARK_DOUBLES + cvx::PRE_DOUBLES) +
                                  slots * cvx::HAND_DOUBLES);
    solve_sdp_kernel<<<(unsigned)blocks, NT, SMEM_BYTES, (cudaStream_t)stream>>>(dd, make_opts(d), Q, hist, slots);
    g_launches = 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    return 0;
}

int cvxpnpl_b200_extract(const cvxpnpl_b200_desc* d, const double* Z, const double* Q, const double* Bmat,
                         const double* dobj, void* stream)
{
    g_launches = 0;
    if (int rc = check_common(d)) return rc;
    if (d->batch == 0) return 0;
    if (!Z || !Q || !Bmat) return fail(-5, "null input pointer");
    if (!d->R || !d->t || !d->n_poses || !d->status) return fail(-6, "null output pointer");
    const size_t smem = (size_t)NT_X * 227 * sizeof(double);
    {
        const cudaError_t e = opt_in_once(3, [smem] {
            return cudaFuncSetAttribute(extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        });
        if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    }
    const int64_t blocks = (d->batch + NT_X - 1) / NT_X;
    const double eps = d->eps > 0 ? d->eps : 1e-9;
    extract_kernel<<<(unsigned)blocks, NT_X, smem, (cudaStream_t)stream>>>(*d, Z, Q, Bmat, dobj, eps);
    g_launches = 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
