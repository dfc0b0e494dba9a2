/*
 * cvxpnpl_b200.h -- C ABI of the B200-native batched CvxPnPL solver.
 *
 * Drop-in boundary for the pose-solve path of SergioRAgostinho/cvxpnpl
 * (reference commit e20cca87).  Every entry point takes plain pointers and
 * sizes; all buffers are DEVICE pointers owned by the caller (the Python host
 * side allocates them as torch tensors); the library allocates nothing that is
 * visible to the user and runs asynchronously on the supplied cudaStream_t.
 *
 * Per-problem array layouts are exactly the reference's, with a leading batch
 * dimension, row-major contiguous, float64:
 *     pts_2d  [B, n_pts, 2]        cvxpnpl.py:28   (pixels)
 *     pts_3d  [B, n_pts, 3]        cvxpnpl.py:29
 *     line_2d [B, n_lines, 2, 2]   cvxpnpl.py:113  (line, endpoint, xy)
 *     line_3d [B, n_lines, 2, 3]   cvxpnpl.py:115  (line, endpoint, xyz)
 *     K       [3, 3] shared, or [B, 3, 3] when k_batched != 0   cvxpnpl.py:30
 *
 * Reference interfaces replaced (file:line in /root/reference):
 *     cvxpnpl_b200_solve     <- pnp  cvxpnpl.py:523-552, pnl 555-583, pnpl 586-627
 *                               (= _point_constraints 20-104, _line_constraints
 *                               107-153, scs.solve 485-489, extraction 493-520)
 *     cvxpnpl_b200_assemble  <- the B / A / Q lines: 548-549, 579-580, 623-624, 475
 *     cvxpnpl_b200_solve_sdp <- scs.solve(...) call, 478-492
 *     cvxpnpl_b200_extract   <- 493-520 with _constraint_ortho_det 221-343 and
 *                               _re6q3 156-218
 * The ctypes binding a maintainer of the reference would add is in INTEGRATION.md.
 */
#ifndef CVXPNPL_B200_H
#define CVXPNPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-problem status (low byte) and flags, written to `status` */
#define CVXPNPL_ST_OK 0             /* converged to eps */
#define CVXPNPL_ST_MAX_ITERS 1      /* iteration cap hit, last iterate used (SCS: solved_inaccurate) */
#define CVXPNPL_ST_NAN 2            /* non-finite data / iterate: one NaN pose (cvxpnpl.py:493-498) */
#define CVXPNPL_ST_SINGULAR 3       /* LinAlgError of cvxpnpl.py:165 / 212 */
#define CVXPNPL_ST_RANK0 4          /* NotImplementedError of cvxpnpl.py:341 */
#define CVXPNPL_ST_CODE_MASK 0xff
#define CVXPNPL_FLAG_NOT_CERTIFIED 0x100 /* warning of cvxpnpl.py:516-519 */

#define CVXPNPL_MAX_POSES 4

typedef struct cvxpnpl_b200_desc {
    /* ---- problem ---- */
    int64_t batch;          /* B */
    int32_t n_pts;          /* points per problem (may be 0) */
    int32_t n_lines;        /* lines per problem (may be 0) */
    int32_t k_batched;      /* 0: K is [3,3]; else [B,3,3] */
    int32_t anderson;       /* Anderson acceleration of the ADMM iteration: 0 = default (on), -1 = off */
    const double* K;
    const double* pts_2d;
    const double* pts_3d;
    const double* line_2d;
    const double* line_3d;
    /* ---- solver options (reference kwargs: cvxpnpl.py:527-529) ---- */
    double eps;             /* fixed-point residual tolerance; reference default 1e-9 */
    int32_t max_iters;      /* reference default 2500 */
    int32_t sweeps;         /* Jacobi sweeps per ADMM iteration (warm started); 0 = default */
    int32_t variant;        /* 0: the reference's SDP (cvxpnpl.py:387-448); 1: "rc" ablation with the six
                               row-orthonormality equalities removed (benchmarks/toolkit/methods/rc.py:9-60) */
    int32_t handoff;        /* stragglers: once the work queue of the persistent kernel is empty, a problem still
                               iterating this many passes later is finished by the warp-per-problem kernel.
                               0 = default (40), < 0 = never hand over */
    double rho_rel;         /* ADMM penalty = rho_rel * ||Q||_F ; 0 = default */
    double alpha;           /* over-relaxation in (0,2); 0 = default */
    double sigma;           /* homogeneous-coordinate scaling (preconditioner); 0 = default */
    /* ---- outputs ---- */
    double* R;              /* [B, 4, 3, 3] world->camera rotations, NaN padded */
    double* t;              /* [B, 4, 3] */
    int32_t* n_poses;       /* [B] 1, 2 or 4 (0 on ST_SINGULAR / ST_RANK0) */
    int32_t* status;        /* [B] */
    int32_t* iters;         /* [B] ADMM iterations used */
    double* obj;            /* optional [B, 2]: r'Qr of candidate 0, dual objective */
    double* Z;              /* optional [B, 10, 10] final PSD iterate */
    /* ---- scratch ---- */
    double* workspace;      /* cvxpnpl_b200_workspace_bytes(batch) bytes */
    size_t workspace_bytes;
    /* ---- precision ---- */
    int32_t fp32_iters;     /* 0: FP64 ADMM throughout (BASELINE.json configs[1], [2]).  > 0: "fp32 ADMM + fp64
                               extraction" (configs[3]): the iterations that bring a problem into the linear
                               tail run in FP32 (at most this many), the FP64 solver finishes to `eps` */
    int32_t timing;         /* != 0: record CUDA events between the kernels of `solve` (cvxpnpl_b200_kernel_times) */
    int32_t skip_prepass;   /* != 0: the pre-pass of every problem has been run by cvxpnpl_b200_prepass already */
    int32_t psd_mode;       /* PSD projection of the ADMM iteration.  0 = default: two tracked eigenpairs refined once per
                               iteration, with a Cholesky certificate that nothing else in the spectrum is positive (problems
                               that fail it for good are finished with the full decomposition), two threads per problem;
                               1 = full 10x10 eigen-decomposition (warm-started Jacobi sweep) every iteration, one thread
                               per problem (the round-1 solver; in `assemble`: the plain-load large-n kernel);
                               2 = tracked eigenpairs, one thread per problem */
    /* ---- packed output (optional) ---- */
    double* record;         /* optional [B, 15]: R of candidate 0 (9, row-major) | t (3) | n_poses | status | iters,
                               written by the finish kernel next to R / t: the row a multi-GPU caller all-gathers
                               (the path's only collective) or copies back to the host, without a packing pass */
} cvxpnpl_b200_desc;

#define CVXPNPL_RECORD 15

/* library version string, e.g. "cvxpnpl_b200 0.1.0 (sm_100a)" */
const char* cvxpnpl_b200_version(void);

/* text of the last error returned on this host thread */
const char* cvxpnpl_b200_last_error(void);

/* bytes of device scratch `solve` / `solve_sdp` need for a batch of B */
size_t cvxpnpl_b200_workspace_bytes(int64_t batch);

/* Full path: assembly -> SDP -> extraction for B problems.  Returns 0 on
 * success, a negative value for bad arguments, a positive cudaError_t otherwise.
 * `stream` is a cudaStream_t.  Asynchronous. */
int cvxpnpl_b200_solve(const cvxpnpl_b200_desc* desc, void* stream);

/* Pre-pass (assembly + start decomposition) of problems [first, first + count) only, so a caller that
 * streams the correspondences from the host can run it chunk by chunk under the copies; call it for
 * every problem of the batch (the chunk with first == 0 first: it resets the workspace header), then
 * `solve` with desc.skip_prepass = 1.  Same descriptor, same workspace. */
int cvxpnpl_b200_prepass(const cvxpnpl_b200_desc* desc, int64_t first, int64_t count, void* stream);

/* Device time of each kernel of the last `solve` issued by this host thread with
 * desc.timing != 0, in ms: pre, admm32, ortho, solve_fused, straggler, resume, finish, solve_track,
 * redecomp (n >= 9; 0 for kernels that were not launched).  Synchronises on the last event. */
int cvxpnpl_b200_kernel_times(float* ms, int n);

/* Stage: correspondences -> Q [B,9,9] (= A'A of cvxpnpl.py:475) and Bmat [B,3,9]
 * (cvxpnpl.py:623).  Uses the problem fields of desc (and psd_mode, see there).  From 256 correspondences per problem
 * the assembly is a streaming reduction; point slabs that are 16-byte aligned (even point count) are staged through
 * shared memory with TMA bulk copies. */
int cvxpnpl_b200_assemble(const cvxpnpl_b200_desc* desc, double* Q, double* Bmat, void* stream);

/* Stage: Q [B,9,9] -> Z [B,10,10] (desc->Z), iterations (desc->iters), status,
 * obj[:,1] = dual objective.  Uses batch, options, workspace. */
int cvxpnpl_b200_solve_sdp(const cvxpnpl_b200_desc* desc, const double* Q, void* stream);

/* Stage: Z [B,10,10], Q [B,9,9], Bmat [B,3,9] -> R, t, n_poses, status.
 * `dobj` ([B], optional) enables the optimality flag. */
int cvxpnpl_b200_extract(const cvxpnpl_b200_desc* desc, const double* Z, const double* Q,
                         const double* Bmat, const double* dobj, void* stream);

/* The "null" baseline of benchmarks/toolkit/methods/pnp.py:24-55 (no SDP: smallest
 * right singular vector of A, SVD projection, det sign fix).  Writes candidate 0 of
 * desc->R / desc->t, n_poses = 1, status. */
int cvxpnpl_b200_null(const cvxpnpl_b200_desc* desc, void* stream);

/* Measurement aid (no reference counterpart): launches an FP64 FMA throughput
 * probe; `out` needs (#SMs * 8 * 256) doubles, *flops receives the flop count of
 * the launch.  bench.py times it with CUDA events to obtain the fp64 peak it
 * quotes next to the HBM roofline. */
int cvxpnpl_b200_fp64_probe(double* out, int64_t out_len, int iters, int64_t* flops, void* stream);

/* Number of kernel launches issued by the last call on this host thread
 * (bench.py reports it as gpu_launches). */
int cvxpnpl_b200_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CVXPNPL_B200_H */
