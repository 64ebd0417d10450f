/*
 * smslu.h -- C ABI of libsmslu.so: B200-native sparse LU refactorization + triangular solves,
 * the drop-in for the hot path of johnomotani/SharedMemSparseLU.jl.
 *
 * The reference has no FFI of its own; its boundary is the Julia API in
 * /root/reference/src/SharedMemSparseLU.jl (cited below as src:N).  Each entry point names the
 * reference interface it replaces; the Julia shim that `ccall`s these is
 * sharedmemsparselu.jl_b200/julia/SharedMemSparseLU.jl and is reproduced in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns int: 0 = OK, negative = SMSLU_E_*; no exceptions, no exit().
 *  - the caller owns every array it passes; nothing is retained after the call returns.
 *    The handle owns all device memory and streams until smslu_destroy.
 *  - indices at the ABI are int64 with an explicit index_base (Julia passes 1).
 *  - value/vector pointers (nzval, Rs, x, b) may be HOST or DEVICE pointers; the library
 *    asks the CUDA runtime which.  Host arrays are staged through pinned buffers.
 *  - a handle is NOT thread-safe (the reference's F.wrk is shared state too, src:52, src:318);
 *    distinct handles are independent.  Calls are synchronous on return.
 *  - there is no CPU fallback: numeric calls fail with SMSLU_E_CUDA when no sm_100 GPU is usable.
 */
#ifndef SMSLU_H
#define SMSLU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMSLU_OK 0
#define SMSLU_E_DIM (-1)      /* -> Julia DimensionMismatch (src:288-290)                        */
#define SMSLU_E_PIVOT (-2)    /* zero / non-finite pivot under the static pivot order -> SingularException */
#define SMSLU_E_PATTERN (-3)  /* malformed CSC, or values do not match the analysed pattern -> ArgumentError */
#define SMSLU_E_ARG (-4)      /* bad argument / wrong call order                                 */
#define SMSLU_E_CUDA (-5)     /* CUDA runtime error or no usable device                           */
#define SMSLU_E_OOM (-6)
#define SMSLU_E_INTERNAL (-7)
#define SMSLU_E_NCCL (-8)
#define SMSLU_E_REPIVOT (-9)  /* the factorization finished, but a multiplier |l_ij| exceeded 1/pivot_tol: the static pivot
                                 order fails UMFPACK's threshold test for these values.  The reference's lu! would re-pivot
                                 (src:245-279, pattern may change: src:252-273); here the caller re-analyses with fresh
                                 (p, q) -- both host mirrors do that automatically.  The factors stay usable.          */

#define SMSLU_ORD_AUTO 0      /* ND_GRID when the grid hint matches n, else ND_GRAPH              */
#define SMSLU_ORD_NATURAL 1
#define SMSLU_ORD_GIVEN 2     /* use the (p,q) handed to smslu_analyze (e.g. UMFPACK's, from the Julia host) */
#define SMSLU_ORD_ND_GRAPH 3  /* level-structure nested dissection on graph(A+A')                 */
#define SMSLU_ORD_ND_GRID 4   /* geometric nested dissection, needs grid[]                         */

#define SMSLU_SCALE_NONE 0
#define SMSLU_SCALE_SUM 1     /* UMFPACK default: Rs[i] = 1/sum_j|a_ij| (what lu(A) at src:74 applies) */

typedef struct smslu_handle_s* smslu_handle_t;

typedef struct smslu_options {
    int32_t ordering;        /* SMSLU_ORD_*                                                   */
    int32_t grid[3];         /* nx, ny, nz with idx = i + nx*(j + ny*k); 0 = unknown          */
    int32_t nd_leaf;         /* stop dissecting below this many vertices                       */
    int32_t relax;           /* relaxed supernode amalgamation on/off                          */
    int32_t max_width;       /* pivot-block width of a front (<= 128); wider supernodes are chained */
    int32_t scaling;         /* SMSLU_SCALE_*, used when smslu_refactor gets Rs == NULL        */
    int32_t device;          /* CUDA device ordinal; -1 = current device                       */
    int32_t reserved0;       /* unused (keeps the layout of version 1.0)                        */
    int32_t nranks;          /* GPUs (= processes) the elimination tree is partitioned over; 0/1 = one GPU */
    int32_t rank;            /* this process' rank in [0, nranks)                                */
    double pivot_tol;        /* threshold test of the static pivots: SMSLU_E_REPIVOT when some |l_ij| > 1/pivot_tol, i.e.
                                |u_jj| < pivot_tol * max_i |a_ij| in column j.  0 = default 1e-3 (UMFPACK's tolerance for
                                diagonal pivots in its symmetric strategy); negative = no test                       */
    int32_t reserved[4];
} smslu_options_t;

typedef struct smslu_stats {
    int64_t n, nnz_a;
    int64_t nnz_l_exact, nnz_u_exact;   /* structural nnz incl. L's unit diagonal                */
    int64_t nnz_l_stored, nnz_u_stored; /* entries held by the supernodal panels per factor      */
    int64_t n_supernodes, n_levels, max_front, max_pivot_block, max_children, sum_rows;
    int64_t lu_pool_doubles, cb_pool_doubles;
    double flops_exact, flops_stored;   /* refactorization flops                                  */
    double ms_analyze, ms_upload;
    double ms_refactor, ms_solve;       /* device time of the last call (CUDA events)             */
    double ms_refactor_h2d, ms_solve_h2d, ms_solve_d2h;
    int64_t launches_refactor, launches_solve;  /* kernels launched by the last call             */
    int64_t n_refactor, n_solve;        /* calls so far                                           */
    int64_t bad_pivot_col;              /* permuted column of the first bad pivot, or -1          */
    double ms_kernel[16];               /* smslu_set_profile(h,1): device ms per kernel kind, summed */
    int64_t launches_kernel[16];        /*   and launches per kernel kind (index = SMSLU_K_*)     */
    int64_t n_top_supernodes;           /* partition: supernodes factored redundantly by every rank      */
    int64_t n_local_supernodes;         /*   supernodes owned by this rank                                */
    int64_t allreduce_doubles_refactor; /*   doubles summed across ranks per refactorization / per solve  */
    int64_t allreduce_doubles_solve;
    int64_t threshold_col;              /* first permuted column of a front whose multipliers failed the threshold test, or -1 */
    int64_t reserved[3];
} smslu_stats_t;

/* kernel kinds for ms_kernel / launches_kernel */
#define SMSLU_K_ROWSCALE 0
#define SMSLU_K_SCATTER 1    /* zero-fill of the big fronts' panels + scatter of their entries of A */
#define SMSLU_K_ZERO 2
#define SMSLU_K_EXTEND 3
#define SMSLU_K_SMALL 4      /* fronts assembled, factored and stored from shared memory */
#define SMSLU_K_PANEL 5
#define SMSLU_K_GEMM 6
#define SMSLU_K_PERMUTE 7
#define SMSLU_K_FWD 8
#define SMSLU_K_BWD 9
#define SMSLU_K_UNPERMUTE 10
#define SMSLU_K_FWD_SMALL 11 /* warp-per-front solve kernels of the shared-memory-sized fronts */
#define SMSLU_K_BWD_SMALL 12
#define SMSLU_K_ALLREDUCE 13 /* NCCL all-reduces of a partitioned handle (+ the small kernels around them) */

/* Fill *opts with defaults.  */
int smslu_options_default(smslu_options_t* opts);

/* Start of `ParallelSparseLU(A::SparseMatrixCSC{Float64,Int64}, chunk_size)` (src:64): take the
 * pattern of A (colptr[n+1], rowval[nnz]).  Host-only; does not touch the GPU. */
int smslu_create(smslu_handle_t* h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                 int32_t index_base, const smslu_options_t* opts);

/* Symbolic half of `lu(A)` (src:74).  p, q: row/column permutations with
 * L*U == (Rs .* A)[p,q] (src:307), index_base-based, only read when ordering == SMSLU_ORD_GIVEN;
 * pass NULL for the native orderings.  Host-only.  The layout is uploaded on first numeric call. */
int smslu_analyze(smslu_handle_t h, const int64_t* p, const int64_t* q);

/* Numeric half of `lu(A)` (src:74) and all of `lu!(F, A)` (src:245-279) for a fixed pattern:
 * nzval[nnz] in the order of the pattern given to smslu_create.  Rs: row multipliers indexed by
 * original row (as UMFPACK's `.Rs`, src:263), or NULL to use opts.scaling. */
int smslu_refactor(smslu_handle_t h, const double* nzval, const double* Rs);

/* `ldiv!(x, F, b)` (src:286-342): x = A \ b; b is not modified.  nrhs = 1 is the reference
 * signature; nrhs > 1 solves column-major blocks with leading dimensions ldx, ldb (BASELINE config 5):
 * 32 columns per sweep on the FP64 tensor pipe while at least 9 are left (one GPU), then 8 / 4 / 1.
 * nx / nb are the lengths of x and b per column (checked against n: SMSLU_E_DIM, src:288-290). */
int smslu_solve(smslu_handle_t h, double* x, int64_t nx, const double* b, int64_t nb,
                int64_t nrhs, int64_t ldx, int64_t ldb);

/* Stream-ordered variants for callers that keep everything on the device (CUDA.jl arrays, the
 * benchmark's kernel-only leg): DEVICE pointers only, work is enqueued on the handle's stream and
 * the call returns without synchronizing.  smslu_sync waits and reports a deferred pivot failure.
 * smslu_set_stream makes the handle use the caller's cudaStream_t (so the caller's events bracket
 * the work); smslu_set_profile turns per-launch event timing on (stats.ms_kernel). */
int smslu_refactor_async(smslu_handle_t h, const double* nzval_dev, const double* Rs_dev);
int smslu_solve_async(smslu_handle_t h, double* x_dev, const double* b_dev);
int smslu_sync(smslu_handle_t h);
int smslu_set_stream(smslu_handle_t h, void* cuda_stream);
int smslu_set_profile(smslu_handle_t h, int32_t on);

/* `lsolve!(F, x)` (src:349-367) and `rsolve!(F, x)` (src:374-392): in place, in the permuted and
 * scaled index space, x <- L^{-1} x  and  x <- U^{-1} x. */
int smslu_lsolve(smslu_handle_t h, double* x, int64_t nx, int64_t nrhs, int64_t ld);
int smslu_rsolve(smslu_handle_t h, double* x, int64_t nx, int64_t nrhs, int64_t ld);

/* Fields `F.L F.U F.p F.q F.Rs` (src:47-51): CSC, rows sorted, L with explicit unit diagonal,
 * exact structural pattern (relaxation padding removed).  Any pointer may be NULL. */
int smslu_get_nnz(smslu_handle_t h, int64_t* nnz_l, int64_t* nnz_u);
int smslu_get_factors(smslu_handle_t h, int64_t* l_colptr, int64_t* l_rowval, double* l_nzval,
                      int64_t* u_colptr, int64_t* u_rowval, double* u_nzval,
                      int64_t* p, int64_t* q, double* Rs, int32_t index_base);

/* Multi-GPU, one process per GPU (no reference counterpart: its MPI dependency is declared but never
 * used, Project.toml:8).  Every rank creates a handle for the SAME pattern with options.nranks /
 * options.rank set; smslu_analyze partitions the elimination tree into per-rank subtrees plus a top part
 * that every rank factors after the subtrees' Schur-complement contributions have been summed with
 * ncclAllReduce over NVLink.  Rank 0 calls smslu_comm_unique_id and hands the 128 bytes to the other
 * ranks (MPI.Bcast, torch.distributed.broadcast, ...); then every rank calls smslu_comm_init
 * (collective).  After that smslu_refactor* / smslu_solve* are collective calls: every rank passes the
 * full nzval / b and receives the full x. */
int smslu_comm_unique_id(void* id, int64_t nbytes);
int smslu_comm_init(smslu_handle_t h, const void* id, int64_t nbytes);

/* Diagnostics (no reference counterpart). */
int smslu_get_stats(smslu_handle_t h, smslu_stats_t* stats);
const char* smslu_last_error(smslu_handle_t h);
/* Symbolic layout read-back for tests: any pointer may be NULL.  sn_start[nsn+1], rows_ptr[nsn+1],
 * rows[sum_rows], sn_parent[nsn], sn_level[nsn], etree_parent[n], colcount[n]. */
int smslu_get_symbolic(smslu_handle_t h, int64_t* sn_start, int64_t* rows_ptr, int64_t* rows,
                       int64_t* sn_parent, int64_t* sn_level, int64_t* etree_parent, int64_t* colcount);

/* `cleanup_ParallelSparseLU!(F)` (exported but undefined in the reference, src:31). */
int smslu_destroy(smslu_handle_t h);

/* `allocate_shared` (exported but undefined in the reference, src:31): name kept, no-op. */
int smslu_allocate_shared(void);

/* Page-locked host buffers for x / b / nzval so host<->device copies run at full PCIe speed
 * (optional: any host pointer is accepted by the numeric calls). */
int smslu_host_alloc(void** ptr, int64_t bytes);
int smslu_host_free(void* ptr);

/* Developer aid: phase timestamps (clock64) of the most recent panel kernel, 32 slots; returns the
 * number of slots filled -- 0 unless the library was built with SMSLU_TRACE=1. */
int smslu_debug_trace(int64_t* out32);

/* Library/ABI version: major*10000 + minor*100 + patch. */
int smslu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SMSLU_H */
