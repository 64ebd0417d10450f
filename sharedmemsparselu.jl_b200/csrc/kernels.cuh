// Hand-written sm_100a kernels for the numeric hot path: multifrontal supernodal LU
// refactorization (what UMFPACK does inside `lu!`, reference src/SharedMemSparseLU.jl:247) and
// the level-scheduled supernodal triangular solves (reference lsolve!/rsolve!, src:349-392).
//
// Data layout in HBM (built once by symbolic.cpp):
//   per supernode s with k pivot columns, r off-diagonal rows, f = k + r:
//     P_s = lu + Loff[s] : f x k column-major (ld f).  Rows [0,k) are the pivot block (after
//           factorization: L11 strictly below the diagonal, U11 on/above); rows [k,f) are L21.
//     T_s = lu + Uoff[s] : r x k column-major (ld r) = U12 transposed.
//     C_s = cb + CBoff[s]: r x r column-major contribution block (temporary).
//   rows[rows_ptr[s] ..] : the r global (permuted) row indices, ascending.
//   rel [rows_ptr[s] ..] : index of each of those rows inside the PARENT's front.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smslu {

constexpr int NB = 32;            // panel block: pivot columns eliminated per panel step
constexpr int KW = 128;           // widest pivot block of a front (symbolic chains wider supernodes)
constexpr int KMAX = KW;
constexpr int PANEL_ROWS = 128;   // rows of L21 / U12' handled by one panel CTA (levels with many fronts)
constexpr int PANEL_ROWS_TOP = 32; // ... near the top of the tree, where one CTA's latency is what matters
constexpr int GEMM_TILE = 64;
constexpr int SOLVE_THREADS = 256;
constexpr int FWD_ROWS = 64;      // update rows per forward CTA (4 threads per row) ...
constexpr int BWD_WIDE_TILES = 148; // backward launches with more tiles than this use the two-CTAs-per-SM variant
constexpr int FWD_ROWS_MID = 128;  // ... two threads per row where 64-row tiles would be more than one wave but 256-row tiles leave SMs idle
constexpr int FWD_ROWS_WIDE = 256; // ... or one thread per row on levels with many tiles (every CTA repeats the pivot-block solve)
constexpr int BWD_ROWS = 256;     // rows of U12' per backward CTA
constexpr int ZERO_TILE = 8192;
constexpr int RB_MAX = 8;         // most right-hand sides swept together by the FMA solve kernels
constexpr int RB_WIDE = 32;       // right-hand sides of a sweep of the tensor-pipe solve kernels (k_fwd32 / k_bwd32)
constexpr int RB_WIDE_ROWS = 256; // rows of L21 / U12' per CTA of those kernels ...
constexpr int RB_WIDE_ROWS_TOP = 64; // ... forward, on levels with few tiles
constexpr int ASM_COLS = 8;       // destination columns of a parent front per assembly CTA
constexpr int MAX_RANKS = 8;      // GPUs of one NVSwitch box
constexpr int CHAIN_MAXOWN = 8;   // blocks (forward) / links (backward) of a chain one CTA of the persistent solve kernels may own
constexpr int CHAIN_DESC = 8;     // ints per chain descriptor: links_off, m, boff_off, nblocks, cta0, nctas, part_off (in KW-vectors), tiles above
constexpr int CHAIN_CTAS = 148;   // one CTA per SM, all co-resident (cooperative launch)
#ifndef SMSLU_ASM_ROWS
#define SMSLU_ASM_ROWS 512
#endif
constexpr int ASM_SMEM_ROWS = SMSLU_ASM_ROWS; // rows of a parent front one assembly CTA holds in shared memory (96 KB per CTA at most; a multiple of 16)
// rows per chunk / chunks of a parent with f rows (taller parents are cut into equal row chunks, one assembly task each)
__host__ __device__ inline int asm_chunks(int64_t f) { return (int)((f + ASM_SMEM_ROWS - 1) / ASM_SMEM_ROWS); }
__host__ __device__ inline int asm_chunk_rows(int64_t f) {
    const int nc = asm_chunks(f);
    return nc <= 1 ? (int)f : (int)((((f + nc - 1) / nc) + 15) & ~(int64_t)15);
}

struct DevCtx {
    const int* sn_start;
    const int64_t* rows_ptr;
    const int* rows;
    const int* rel;
    const int64_t* Loff;
    const int64_t* Uoff;
    const int64_t* CBoff;
    const int* sn_parent;
    const int* child_ptr;
    const int* child_idx;
    double* lu;
    double* cb;
    double* upd;        // forward-solve update vectors, sum_r doubles
    double* bpart;      // backward-solve partial sums, KMAX doubles per (supernode, tile)
    int* counters;      // one per (big front, panel step), used by the panel kernel
    int* counters2;     // one per supernode, used by the backward-solve kernel
    int* flag;          // flag[0]: first bad (zero / non-finite) pivot column, flag[1]: first column of a front with a
                        // multiplier above lmax (threshold test); atomicMin, 0x7f7f7f7f when clean
    double lmax;        // 1 / pivot_tol (infinity: no test)
    double* dinv;       // 1 / u_jj by permuted column, written by the factor kernels
    const int* Doff;    // big fronts: index of the front's first 32x32 block in dblk
    double* dblk;       // inverses of the 32x32 diagonal blocks of the big fronts' pivot blocks (k_diag_inverse)
    // entries of A grouped by the small front that pulls them (k_small_factor)
    const int* a_ptr;   // nsn+1 (empty range for big fronts)
    const int* a_src;   // index into the caller's nzval
    const int* a_row;   // original row (for Rs)
    const int* a_pos;   // row | col << 16 inside the front
    // assembly kernel: per task, the (child, first column in range) pairs it pulls
    const int* asm_meta;
    // partition over several GPUs (one process each), see DESIGN.md "Multi-GPU": owner of the global column behind
    // every entry of `rows` (-1: not a top column); this rank; the ranks' contribution / factor pools and sync flags
    // mapped into this process (CUDA IPC; own pool at index `rank`)
    // chains of fronts cut out of one wide separator, solved by one persistent kernel per sweep (k_fwd_chain / k_bwd_chain)
    const int* chain_links;   // supernodes of every chain, bottom link first
    const int* chain_boff;    // per chain: row offsets of its blocks (pivot blocks of the links, then 128-row blocks of the rows above the chain)
    int* chain_flags;         // [2 * chain_nsn]: epoch at which link s published its part of the solution (forward | backward)
    int chain_nsn;
    double* chain_part;       // backward: partial products of the rows above the chain, [link][row tile][KW]
    const signed char* rowown;
    int rank, nranks;
    double* cb_peer[MAX_RANKS];
    double* lu_peer[MAX_RANKS];
    int* xflag_peer[MAX_RANKS];
};

// ---- refactorization
void launch_rowscale(cudaStream_t st, int n, const int64_t* rowptr, const int64_t* rowidx, const double* av, double* Rs);
void launch_scatter(cudaStream_t st, int64_t nnz, const int64_t* dst, const int* arow, const int* asrc,
                    const double* Rs, const double* av, double* lu);
void launch_zero_cb(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks);
void launch_diag_inverse(cudaStream_t st, const DevCtx& cx, const int2* tasks, int ntasks);
void launch_assemble(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int fmax);
void launch_front_small(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int fmax,
                        const double* av, const double* Rs);
// rb = right-hand sides swept together (1, 4 or 8); vectors are interleaved [i * rb + q]
// ---- partitioned top of the tree (nranks > 1)
// copy factor-pool segments (tasks: x = offset / 2, y = length / 2 in double2 units ... as int64 pairs) to every peer
void launch_replicate(cudaStream_t st, const DevCtx& cx, const int64_t* segs, int nsegs);
// rows of U12' this rank owns, of the top fronts listed in tasks (x = supernode, y = first row, z = rows), to every peer
void launch_replicate_rows(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks);
void launch_signal(cudaStream_t st, const DevCtx& cx, int slot, int epoch);
void launch_wait(cudaStream_t st, const DevCtx& cx, const int* slots, int nslots, int epoch);
void launch_vgather(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const int* vlist, int rb);
void launch_mask_owned(cudaStream_t st, int n, const int* colowner, int rank, double* z, int rb);
void launch_small_fwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const double* win, double* zout, int rb);
void launch_small_bwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x, int rb);
void launch_panel(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int g, int rows);
void launch_gemm_cb(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks);
void launch_gemm_strip(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks);   // y = first row tile | tiles << 16
int front_small_limit();   // largest front the fused shared-memory kernel takes
cudaError_t kernels_init();
int debug_read_trace(long long* out);   // 0 unless built with SMSLU_TRACE

// ---- solves
void launch_permute_scale(cudaStream_t st, int n, const int* p, const double* Rs, const double* b, int64_t ldb, double* w, int rb, int nv);
void launch_unpermute(cudaStream_t st, int n, const int* q, const double* w, double* x, int64_t ldx, int rb, int nv);
void launch_fwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int rows, const double* win, double* zout, int rb);
void launch_bwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x, int rb);
// 32 right-hand sides per sweep on the FP64 tensor pipe (tasks: x = supernode, y = row tile of RB_WIDE_ROWS rows; backward also
// z = tiles of the supernode, w = first slot of its partial sums)
void launch_fwd32(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int rows, const double* win, double* zout);
void launch_bwd32(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x);
// persistent chain kernels (single right-hand side): descriptors of `nchains` chains, `nctas` CTAs in total
cudaError_t launch_fwd_chain(cudaStream_t st, const DevCtx& cx, const int* chains, int nchains, int nctas, const double* win, double* zout, int epoch);
void launch_bwd_rect(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const double* x);
cudaError_t launch_bwd_chain(cudaStream_t st, const DevCtx& cx, const int* chains, int nchains, int nctas, double* x, int epoch);

}  // namespace smslu
