// sm_100a kernels, see kernels.cuh for the data layout.  FP64 throughout.
#include "kernels.cuh"

#include <math.h>
#include <stdlib.h>

namespace smslu {

// Optional phase timestamps (build with SMSLU_TRACE=1): CTA 0 of the most recent k_panel launch records
// clock64() at its phase boundaries; slots 0-7 = row warp 0, 8-15 = pivot warp.
#ifdef SMSLU_TRACE
__device__ long long g_trace[32];
#define TRACE(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) == 4)) \
        g_trace[(i) + ((threadIdx.x >> 5) == 4 ? 8 : 0)] = clock64(); } while (0)
#define TRACE2(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && gridDim.x == 1) g_trace[16 + (i)] = clock64(); } while (0)
// panel kernel: SMSLU_TRACE_ROWS selects the variant that records (default: the 32-row latency variant)
#ifndef SMSLU_TRACE_ROWS
#define SMSLU_TRACE_ROWS 32
#endif
#define TRACEP(i) do { if (ROWS == SMSLU_TRACE_ROWS) TRACE(i); } while (0)
#else
#define TRACE(i) do {} while (0)
#define TRACE2(i) do {} while (0)
#define TRACEP(i) do {} while (0)
#endif

namespace {

struct Front {
    int c0, k;
    int64_t r, f;
    double* P;
    double* T;
    double* C;
};

__device__ __forceinline__ Front load_front(const DevCtx& cx, int s) {
    Front F;
    F.c0 = cx.sn_start[s];
    F.k = cx.sn_start[s + 1] - F.c0;
    F.r = cx.rows_ptr[s + 1] - cx.rows_ptr[s];
    F.f = F.k + F.r;
    F.P = cx.lu + cx.Loff[s];
    F.T = cx.lu + cx.Uoff[s];
    F.C = cx.cb + cx.CBoff[s];
    return F;
}

// Programmatic dependent launch: every task kernel lets its successor start launching right away and reads its
// own static metadata (task record, front geometry) before waiting for the predecessor's results, so the
// successor's prologue overlaps the predecessor's tail.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Pull a rows x cols column-major block (leading dimension ld) towards L2 ahead of its use: one prefetch per
// 128 bytes of every column (the block's start need not be line-aligned, hence the clamped extra probe).
// The factors are never written during a solve, so this is issued before griddepcontrol.wait.
__device__ __forceinline__ void prefetch_block_l2(const double* base, int rows, int cols, int64_t ld, int tid, int nthreads) {
    if (rows <= 0) return;
    const int per = (rows + 15) / 16 + 1;
    for (int e = tid; e < cols * per; e += nthreads) {
        const int c = e / per, l = e - c * per;
        const int off = l * 16 < rows ? l * 16 : rows - 1;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)c * ld + off));
    }
}

// Warp index as a provably warp-uniform value (a shuffle from lane 0): branches on it are uniform branches, so the
// compiler does not wrap every shuffle inside a warp-role branch into WARPSYNC.COLLECTIVE / ENDCOLLECTIVE + moves
// (that wrapping tripled the instruction count of the 32 x 32 register LU).
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ bool bad_pivot(double p) { return !(fabs(p) > 0.0) || !isfinite(p); }

constexpr int SMALL_F_MAX = 96;   // largest front handled by the shared-memory kernels (with k <= NB)
__host__ __device__ constexpr int small_ld(int f) { return f + ((6 - (f & 3)) & 3); }   // smallest ld >= f with ld = 2 (mod 4)
__host__ __device__ constexpr int small_group_doubles(int fmax) { return fmax * small_ld(fmax) + NB; }

// ------------------------------------------------------------------ row scaling (UMFPACK "SUM")
__global__ void k_rowscale(int n, const int64_t* __restrict__ rowptr, const int64_t* __restrict__ rowidx,
                           const double* __restrict__ av, double* __restrict__ Rs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int64_t t = rowptr[i]; t < rowptr[i + 1]; ++t) s += fabs(av[rowidx[t]]);   // ascending column
    Rs[i] = s > 0.0 ? 1.0 / s : 1.0;
}

// ------------------------------------------------------------------ A -> panels
// (entries of the big fronts only; the small fronts pull theirs inside k_small_factor)
__global__ void k_scatter(int64_t nnz, const int64_t* __restrict__ dst, const int* __restrict__ arow,
                          const int* __restrict__ asrc, const double* __restrict__ Rs,
                          const double* __restrict__ av, double* __restrict__ lu) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride)
        lu[dst[t]] = Rs[arow[t]] * av[asrc[t]];
}

// ------------------------------------------------------------------ zero contribution blocks
// task: x = supernode, y = tile
__global__ void __launch_bounds__(256) k_zero_cb(DevCtx cx, const int4* __restrict__ tasks) {
    pdl_trigger();
    int4 tk = tasks[blockIdx.x];
    int64_t r = cx.rows_ptr[tk.x + 1] - cx.rows_ptr[tk.x];
    int64_t tot = r * r;
    double* C = cx.cb + cx.CBoff[tk.x];
    pdl_wait();
    int64_t lo = (int64_t)tk.y * ZERO_TILE, hi = lo + ZERO_TILE;
    if (hi > tot) hi = tot;
    for (int64_t e = lo + threadIdx.x; e < hi; e += 256) C[e] = 0.0;
}

// ------------------------------------------------------------------ assemble children into a big parent
// One CTA owns ASM_COLS destination columns [pb0, pb0 + ncols) of the parent's front and pulls every
// qualifying child's contribution block into them, child by child in ascending order (fixed summation
// order, no atomics, one launch per level).  Destination column pb < k lands in P(:, pb); pb >= k lands
// in row pb-k of U12' (rows pa < k) and in column pb-k of the contribution block (rows pa >= k), which the
// CTA zero-fills first when asked to.  rel is ascending, so a child's columns that fall into the CTA's
// range are found by binary search.
// The host precomputes, per task, the list of (child, first child column that falls into the range):
// the kernel's chain of dependent global loads is task -> child record -> rel -> data.
// task: x = parent, y = pb0, z = offset of the task's (child, lo) pairs in asm_meta, w = zero flag | pairs << 8.
__global__ void __launch_bounds__(256) k_assemble(DevCtx cx, const int4* __restrict__ tasks) {
    const int4 tk = tasks[blockIdx.x];
    const int s = tk.x, pb0 = tk.y, nch = tk.w >> 8;
    pdl_trigger();
    const int* __restrict__ meta = cx.asm_meta + tk.z;
    const Front F = load_front(cx, s);
    pdl_wait();
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wdt = (tk.w >> 4) & 15 ? (tk.w >> 4) & 15 : ASM_COLS;     // narrower where the columns' owner changes
    const int pb1 = pb0 + wdt < (int)F.f ? pb0 + wdt : (int)F.f;
    if (tk.w & 1) {
        const int c_lo = pb0 > k ? pb0 - k : 0, c_hi = pb1 - k;      // contribution-block columns owned here
        if (c_hi > c_lo) {
            double* __restrict__ z = F.C + (int64_t)c_lo * F.r;
            const int64_t cnt = (int64_t)(c_hi - c_lo) * F.r;
            for (int64_t e = tid; e < cnt; e += 256) z[e] = 0.0;
        }
        __syncthreads();
    }
    for (int q = 0; q < nch; ++q) {
        const int c = meta[2 * q], lo = meta[2 * q + 1];
        const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ Cc = cx.cb + cx.CBoff[c];
        for (int b = lo + warp; b < rc; b += 8) {
            const int pb = rel[b];
            if (pb >= pb1) break;
            const double* __restrict__ src = Cc + (int64_t)b * rc;
            // destination of row position pa: P(:, pb) for pb < k; else row pb-k of U12' when pa < k,
            // column pb-k of the contribution block otherwise.  rel is strictly ascending, so the four
            // read-modify-writes of a batch never alias: their loads are issued together.
            double* __restrict__ dP = F.P + (int64_t)pb * F.f;
            double* __restrict__ dC = F.C + (int64_t)(pb - k) * F.r - k;
            double* __restrict__ dT = F.T + (pb - k);
            for (int a0 = lane; a0 < rc; a0 += 128) {
                double* d[4]; double v[4], o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int a = a0 + 32 * u;
                    d[u] = nullptr; v[u] = 0.0;
                    if (a < rc) {
                        const int pa = rel[a];
                        v[u] = src[a];
                        d[u] = pb < k ? dP + pa : (pa < k ? dT + (int64_t)pa * F.r : dC + pa);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) o[u] = d[u] ? *d[u] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (d[u]) *d[u] = o[u] + v[u];
            }
        }
        __syncthreads();
    }
}

// Same task format (y = pb0 | row chunk << 20), the production path: the CTA's destination columns are accumulated in
// shared memory -- loaded once (zeros for the part of the contribution block that would have been zero-filled), the
// children added in the same ascending order (so the result is bitwise the same as k_assemble's), written back once.
// The per-child work has no dependent global round trip left (k_assemble: rel -> old value -> store, per child, 69 % of
// its cycles on the long scoreboard) and the parent is read and written once instead of once per child.
// Parents with more than ASM_SMEM_ROWS rows are cut into row chunks of equal size (asm_chunk_rows), one task per
// (column range, chunk): rel is ascending, so the rows of a child that fall into a chunk are one range, found by two
// binary searches per child (one thread each, beside the loads of the destination block).
// dynamic shared memory: ASM_COLS * min(rows of the largest parent of the launch, ASM_SMEM_ROWS) doubles.
#ifndef ASM_MINB
#define ASM_MINB 5
#endif
__global__ void __launch_bounds__(256, ASM_MINB) k_assemble_smem(DevCtx cx, const int4* __restrict__ tasks) {
    extern __shared__ __align__(16) double cols[];         // cols[c * nr + (pa - ra0)]
    __shared__ int s_alo[64], s_ahi[64];
    const int4 tk = tasks[blockIdx.x];
    const int s = tk.x, pb0 = tk.y & 0xfffff, chunk = (int)((unsigned)tk.y >> 20), nch = tk.w >> 8;
    pdl_trigger();
    const int* __restrict__ meta = cx.asm_meta + tk.z;
    const Front F = load_front(cx, s);
    pdl_wait();
    const int k = F.k, f = (int)F.f, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wdt = (tk.w >> 4) & 15 ? (tk.w >> 4) & 15 : ASM_COLS;
    const int pb1 = pb0 + wdt < f ? pb0 + wdt : f;
    const bool zero = tk.w & 1;
    const int cr = asm_chunk_rows(f), ra0 = chunk * cr, ra1 = ra0 + cr < f ? ra0 + cr : f, nr = ra1 - ra0;
    const bool whole = nr == f;
    // destination block -> shared memory, column by column (no index division): column pb of the front is one contiguous
    // run of P (pb < k), or k strided entries of U12' followed by a contiguous run of the contribution block
    const int ncol = pb1 - pb0;
    for (int c = 0; c < ncol; ++c) {
        const int pb = pb0 + c;
        double* __restrict__ dst = cols + c * nr - ra0;
        if (pb < k) {
            const double* __restrict__ src = F.P + (int64_t)pb * F.f;
            for (int pa = ra0 + tid; pa < ra1; pa += 256) dst[pa] = src[pa];
        } else {
            const double* __restrict__ srcT = F.T + (pb - k);
            const double* __restrict__ srcC = F.C + (int64_t)(pb - k) * F.r - k;
            for (int pa = ra0 + tid; pa < ra1; pa += 256)
                dst[pa] = pa < k ? srcT[(int64_t)pa * F.r] : (zero ? 0.0 : srcC[pa]);
        }
    }
    for (int q0 = 0; q0 < nch; q0 += 64) {
        if (!whole && tid < 128) {                               // rows [alo, ahi) of child q fall into [ra0, ra1)
            const int q = q0 + (tid & 63);
            if (q < nch) {
                const int c = meta[2 * q];
                const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
                const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
                const int key = tid < 64 ? ra0 : ra1;
                int lo = 0, hi = rc;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (rel[mid] < key) lo = mid + 1; else hi = mid; }
                (tid < 64 ? s_alo : s_ahi)[tid & 63] = lo;
            }
        }
        __syncthreads();
        const int q1 = q0 + 64 < nch ? q0 + 64 : nch;
        for (int q = q0; q < q1; ++q) {
            const int c = meta[2 * q], lo = meta[2 * q + 1];
            const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
            const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
            const double* __restrict__ Cc = cx.cb + cx.CBoff[c];
            const int alo = whole ? 0 : s_alo[q - q0], ahi = whole ? rc : s_ahi[q - q0];
            for (int b = lo + warp; b < rc; b += 8) {
                const int pb = rel[b];
                if (pb >= pb1) break;
                const double* __restrict__ src = Cc + (int64_t)b * rc;
                double* col = cols + (pb - pb0) * nr - ra0;
                for (int a0 = alo + lane; a0 < ahi; a0 += 256) {           // eight entries per lane in flight
                    int pa[8]; double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int a = a0 + 32 * u;
                        pa[u] = a < ahi ? rel[a] : -1;
                        v[u] = a < ahi ? src[a] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (pa[u] >= 0) col[pa[u]] += v[u];
                }
            }
            __syncthreads();
        }
    }
    if (nch == 0) __syncthreads();
    for (int c = 0; c < ncol; ++c) {
        const int pb = pb0 + c;
        const double* __restrict__ src = cols + c * nr - ra0;
        if (pb < k) {
            double* __restrict__ dst = F.P + (int64_t)pb * F.f;
            for (int pa = ra0 + tid; pa < ra1; pa += 256) dst[pa] = src[pa];
        } else {
            double* __restrict__ dstT = F.T + (pb - k);
            double* __restrict__ dstC = F.C + (int64_t)(pb - k) * F.r - k;
            for (int pa = ra0 + tid; pa < ra1; pa += 256) {
                if (pa < k) dstT[(int64_t)pa * F.r] = src[pa];
                else dstC[pa] = src[pa];
            }
        }
    }
}

// Pull the front's entries of Rs .* A into its shared-memory image, four entries per thread in flight (the index
// loads and the two dependent gathers are latency, not bandwidth).  BEFORE: the loads of the first batch are issued
// ahead of the caller's zero-fill barrier through the two-phase interface below.
struct PulledEntries {
    int pos[4];
    double val[4];
};
__device__ __forceinline__ void pull_entries_load(const DevCtx& cx, int e, int e1, int nt, const double* __restrict__ av,
                                                  const double* __restrict__ Rs, PulledEntries& b) {
    int row[4], src[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int ee = e + u * nt;
        const bool ok = ee < e1;
        b.pos[u] = ok ? cx.a_pos[ee] : -1;
        row[u] = ok ? cx.a_row[ee] : 0;
        src[u] = ok ? cx.a_src[ee] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) b.val[u] = b.pos[u] >= 0 ? Rs[row[u]] * av[src[u]] : 0.0;
}
__device__ __forceinline__ void pull_entries_store(const PulledEntries& b, double* Fs, int ld) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (b.pos[u] >= 0) Fs[(b.pos[u] & 0xffff) * ld + (b.pos[u] >> 16)] = b.val[u];
}

// ------------------------------------------------------------------ fused small front
// A front with k <= 32 pivots and f <= SMALL_F_MAX rows is assembled, factored and stored by one
// group of RW*CH threads entirely in shared memory ("pull" assembly):
//   Fs (f x f, ROW-major, leading dimension ld = 2 mod 4 so that 128-bit row accesses of consecutive
//   lanes are conflict-free) = 0
//   Fs += the entries of Rs .* A that land in this front      (per-front lists a_ptr/a_src/a_row/a_pos)
//   Fs += contribution blocks of the children, child by child in ascending order (rel maps)
//   right-looking elimination of the k pivots, one barrier per pivot: thread (i, h) owns row i and the
//   column pairs p = (j+1)/2 + h (mod CH), updated with 128-bit loads/stores; multipliers are formed as
//   a_ij * (1/u_jj) on the fly and column j is left unscaled until the final copy-out, so no thread ever
//   reads what another thread writes inside a step.  1/u_jj is produced by the thread that finishes u_jj.
//   copy-out: P (f x k), T = U12' (r x k), C (r x r), dinv (reciprocal pivots for the solves).
// Factor storage is written exactly once and never read here; nothing has to be zero-filled.
// RW = row slots (>= f), CH = column interleave, FPC = fronts per CTA (FPC > 1 only with one warp
// per front, where the barriers are __syncwarp).
// task: x = supernode.
template <int RW, int CH, int FPC>
__global__ void __launch_bounds__(RW * CH * FPC) k_small_factor(DevCtx cx, const int4* __restrict__ tasks, int ntasks,
                                                                 int fmax, const double* __restrict__ av,
                                                                 const double* __restrict__ Rs) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = RW * CH;
    static_assert(NT % 32 == 0 && (FPC == 1 || NT == 32), "group layout");
    const int grp = threadIdx.x / NT, tid = threadIdx.x % NT, lane = tid & 31, wrp = tid >> 5;
    const int ti = blockIdx.x * FPC + grp;
    if (ti >= ntasks) return;                                 // FPC > 1: a whole warp leaves
    auto sync = [&]() { if (NT == 32) __syncwarp(); else __syncthreads(); };
    pdl_trigger();
    const int s = tasks[ti].x;
    const Front F = load_front(cx, s);
    pdl_wait();
    const int k = F.k, r = (int)F.r, f = (int)F.f, ld = small_ld(f);
    double* Fs = sm + (size_t)grp * small_group_doubles(fmax);
    double* rd = Fs + (size_t)fmax * small_ld(fmax);
    {
        const int e0 = cx.a_ptr[s] + tid, e1 = cx.a_ptr[s + 1];
        PulledEntries pe;
        pull_entries_load(cx, e0, e1, NT, av, Rs, pe);        // in flight across the zero-fill
        double2* z = reinterpret_cast<double2*>(Fs);
        for (int e = tid; e < f * ld / 2; e += NT) z[e] = make_double2(0.0, 0.0);
        sync();
        pull_entries_store(pe, Fs, ld);
        for (int e = e0 + 4 * NT; e < e1; e += 4 * NT) {
            pull_entries_load(cx, e, e1, NT, av, Rs, pe);
            pull_entries_store(pe, Fs, ld);
        }
    }
    sync();
    for (int ci = cx.child_ptr[s]; ci < cx.child_ptr[s + 1]; ++ci) {
        const int c = cx.child_idx[ci];
        const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ Cc = cx.cb + cx.CBoff[c];
        for (int b = wrp; b < rc; b += NT / 32) {
            const int pb = rel[b];
            for (int a = lane; a < rc; a += 32) Fs[rel[a] * ld + pb] += Cc[a + (int64_t)b * rc];
        }
        sync();
    }
    if (tid == 0) {
        const double piv = Fs[0];
        if (bad_pivot(piv)) atomicMin(cx.flag, F.c0);
        rd[0] = 1.0 / piv;
    }
    sync();
    {
        const int i = tid % RW, h = tid / RW, npair = ld >> 1;
        double* myrow = Fs + i * ld;
        bool big_l = false;                                   // threshold test: some multiplier above 1 / pivot_tol
        for (int j = 0; j < k; ++j) {
            if (i > j && i < f) {
                const double l = myrow[j] * rd[j];
                big_l |= fabs(l) > cx.lmax;
                const double2* __restrict__ prow = reinterpret_cast<const double2*>(Fs + j * ld);
                double2* __restrict__ mrow = reinterpret_cast<double2*>(myrow);
                int p = ((j + 1) >> 1) + h;
                if (h == 0) {                                 // the pair holding column j+1: the next pivot
                    const double2 u = prow[p];
                    double2 v = mrow[p];
                    if (2 * p > j) v.x -= l * u.x;
                    v.y -= l * u.y;
                    mrow[p] = v;
                    if (i == j + 1 && i < k) {
                        const double piv = (j & 1) ? v.x : v.y;
                        if (bad_pivot(piv)) atomicMin(cx.flag, F.c0 + i);
                        rd[i] = 1.0 / piv;
                    }
                    p += CH;
                }
#pragma unroll 2
                for (; p < npair; p += CH) {
                    const double2 u = prow[p];
                    double2 v = mrow[p];
                    v.x -= l * u.x;
                    v.y -= l * u.y;
                    mrow[p] = v;
                }
            }
            sync();
        }
        if (big_l) atomicMin(cx.flag + 1, F.c0);
    }
    for (int c = wrp; c < k; c += NT / 32) {                  // P: pivot block on top of L21
        const double rc = rd[c];
        double* __restrict__ dst = F.P + (int64_t)c * f;
        for (int i = lane; i < f; i += 32) { const double v = Fs[i * ld + c]; dst[i] = i > c ? v * rc : v; }
    }
    for (int c = wrp; c < k; c += NT / 32) {                  // T = U12 transposed
        double* __restrict__ dst = F.T + (int64_t)c * r;
        for (int b = lane; b < r; b += 32) dst[b] = Fs[c * ld + k + b];
    }
    for (int b = wrp; b < r; b += NT / 32) {                  // contribution block for the parent
        double* __restrict__ dst = F.C + (int64_t)b * r;
        for (int a = lane; a < r; a += 32) dst[a] = Fs[(k + a) * ld + k + b];
    }
    if (tid < k) cx.dinv[F.c0 + tid] = rd[tid];
}

// ------------------------------------------------------------------ fused small front, rows in registers
// Fronts with f <= NC (NC = 32, 40, 48, 64): one thread per row, ceil(NC/32) warps per front (FPC fronts
// per CTA when that is one warp).  Assembly in shared memory as in k_small_factor (row-major, ld = NC + 2),
// then every thread takes its row into registers (NC doubles) and the elimination runs without touching
// the front in shared memory again: the owner of row j publishes its (final) row and 1/u_jj in a double-
// buffered strip, every thread reads it back with broadcast 128-bit loads and updates its own registers.
// L21 / the pivot block go to HBM straight from registers (threads = rows: coalesced); U12' and the
// contribution block are transposed through the shared-memory copy of the front.
__host__ __device__ constexpr int reg_group_doubles(int nc) { return nc * (nc + 2) + 2 * (nc + 4) + NB; }

template <int NC, int FPC>
__global__ void __launch_bounds__(((NC + 31) / 32) * 32 * FPC) k_small_factor_reg(DevCtx cx, const int4* __restrict__ tasks, int ntasks,
                                                                                  const double* __restrict__ av,
                                                                                  const double* __restrict__ Rs) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = ((NC + 31) / 32) * 32, ld = NC + 2, SL = NC + 4;
    static_assert(FPC == 1 || NT == 32, "several fronts per CTA only with one warp per front");
    const int grp = threadIdx.x / NT, tid = threadIdx.x % NT;
    const int ti = blockIdx.x * FPC + grp;
    if (ti >= ntasks) return;
    auto sync = [&]() { if (NT == 32) __syncwarp(); else __syncthreads(); };
    pdl_trigger();
    const int s = tasks[ti].x;
    const Front F = load_front(cx, s);
    pdl_wait();
    const int k = F.k, r = (int)F.r, f = (int)F.f;
    double* Fs = sm + (size_t)grp * reg_group_doubles(NC);
    double* strip = Fs + NC * ld;                     // strip[2][SL]: row j (NC values), 1/u_jj at [NC]
    double* rd = strip + 2 * SL;
    {
        const int e0 = cx.a_ptr[s] + tid, e1 = cx.a_ptr[s + 1];
        PulledEntries pe;
        pull_entries_load(cx, e0, e1, NT, av, Rs, pe);        // in flight across the zero-fill
        double2* z = reinterpret_cast<double2*>(Fs);
        for (int e = tid; e < f * ld / 2; e += NT) z[e] = make_double2(0.0, 0.0);
        sync();
        pull_entries_store(pe, Fs, ld);
        for (int e = e0 + 4 * NT; e < e1; e += 4 * NT) {
            pull_entries_load(cx, e, e1, NT, av, Rs, pe);
            pull_entries_store(pe, Fs, ld);
        }
    }
    sync();
    for (int ci = cx.child_ptr[s]; ci < cx.child_ptr[s + 1]; ++ci) {
        const int c = cx.child_idx[ci];
        const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ Cc = cx.cb + cx.CBoff[c];
        const int pa = tid < rc ? rel[tid] * ld : 0;              // rc <= f <= NC <= NT
        for (int b0 = 0; b0 < rc; b0 += 8) {                      // eight columns of the child's block in flight
            double cv[8]; int pb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = tid < rc && b0 + u < rc;
                pb[u] = ok ? rel[b0 + u] : -1;
                cv[u] = ok ? Cc[tid + (b0 + u) * rc] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) if (pb[u] >= 0) Fs[pa + pb[u]] += cv[u];
        }
        sync();
    }
    double x[NC];                                     // row `tid` of the front (zero beyond f)
    bool big_l = false;                               // threshold test: some multiplier above 1 / pivot_tol
#pragma unroll
    for (int p = 0; p < NC / 2; ++p) {
        const double2 v = tid < f ? *reinterpret_cast<const double2*>(Fs + tid * ld + 2 * p) : make_double2(0.0, 0.0);
        x[2 * p] = v.x; x[2 * p + 1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j >= k) break;
        double* sj = strip + (j & 1) * SL;
        if (tid == j) {
            const double piv = x[j];
            if (bad_pivot(piv)) atomicMin(cx.flag, F.c0 + j);
            const double rinv = 1.0 / piv;
            sj[NC] = rinv;
            rd[j] = rinv;
#pragma unroll
            for (int p = (j + 1) / 2; p < NC / 2; ++p) *reinterpret_cast<double2*>(sj + 2 * p) = make_double2(x[2 * p], x[2 * p + 1]);
        }
        sync();
        const double l = tid > j ? x[j] * sj[NC] : 0.0;
        big_l |= fabs(l) > cx.lmax;
        x[j] = tid > j ? l : x[j];
#pragma unroll
        for (int p = (j + 1) / 2; p < NC / 2; ++p) {
            const double2 u = *reinterpret_cast<const double2*>(sj + 2 * p);
            if (2 * p > j) x[2 * p] -= l * u.x;
            x[2 * p + 1] -= l * u.y;
        }
    }
    if (big_l) atomicMin(cx.flag + 1, F.c0);
    // pivot block on top of L21: column c of P, threads = rows (multipliers are already scaled)
    if (tid < f) {
        double* __restrict__ dst = F.P + tid;
#pragma unroll
        for (int c = 0; c < NB; ++c) if (c < k) dst[(int64_t)c * f] = x[c];
    }
    sync();
    if (tid < f) {
#pragma unroll
        for (int p = 0; p < NC / 2; ++p) *reinterpret_cast<double2*>(Fs + tid * ld + 2 * p) = make_double2(x[2 * p], x[2 * p + 1]);
    }
    sync();
    if (tid < r) {
        for (int c = 0; c < k; ++c) F.T[tid + (int64_t)c * r] = Fs[c * ld + k + tid];          // T[b + c r] = U[c][k + b]
        for (int b = 0; b < r; ++b) F.C[tid + (int64_t)b * r] = Fs[(k + tid) * ld + k + b];    // contribution block
    }
    if (tid < k) cx.dinv[F.c0 + tid] = rd[tid];
}

// ------------------------------------------------------------------ FP64 tensor-core tile
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // D(8x8) = A(8x4, row) * B(4x8, col) + C on the FP64 tensor pipe (SASS: DMMA.8x8x4).
    // lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4)], C[l/4][2*(l%4)+1].
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ panel step of a big front
// A front with k (<= KW) pivot columns is factored in ceil(k/32) left-looking panel steps, one
// launch each.  Step g owns pivot columns [j0, j1) = [32g, min(k, 32g+32)):
//   kind 0 (L): ROWS rows of P below the diagonal block (the rest of the pivot block and L21):
//               row <- (row[j0:j1] - row[0:j0] * U[0:j0, j0:j1]) * U_gg^{-1}
//   kind 1 (T): ROWS rows of U12' :  row <- (row[j0:j1] - row[0:j0] * L[j0:j1, 0:j0]') * L_gg^{-T}
//   kind 2 (I): ROWS columns of the pivot block right of the diagonal block (U inside the block),
//               same arithmetic as kind 1 on P[j0:j1, c].
// One CTA = 8 warps, four phases separated by barriers:
//   S  stage D_gg and the two j0 x 32 coefficient blocks in shared memory with cp.async (every element
//      is in flight at once; out-of-range elements are zero-filled by the copy itself);
//   U  left-looking update on the FP64 tensor pipe (DMMA m8n8k4): the ROWS x 32 row block and the
//      32 x 32 diagonal block are cut into pieces of 8 rows x 16 columns that the warps take round-robin
//      (one warp can issue a DMMA only every ~45 cycles, so latency needs many warps, not big tiles);
//      A fragments of the row strips come straight from global memory (one k-step of prefetch);
//   L  warp 0 factors the diagonal block: one row per lane, all 32 columns in registers, pivot row by
//      shuffles, reciprocal pivots; it then publishes the factor rows (W) and, if this CTA was the last
//      of the step to have read the raw block, stores the block and the reciprocal pivots (nobody waits);
//   T  thread = row: solve against the factored block from registers (128-bit loads of W) and store.
// ROWS = 128 when a level has enough fronts to fill the machine, 32 near the top of the tree where the
// latency of a single CTA is what matters.
// task: x = supernode, y = g | kind << 4 | (CTAs of this step of this front) << 8, z = tile,
//       w = index of the step's arrival counter.
// dynamic shared memory: (2 * j0 + ROWS) * CLD doubles.
constexpr int PANEL_THREADS = 256;
constexpr int CLD = NB + 2;           // even row stride: 128-bit aligned pairs

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int nbytes = valid ? 8 : 0;                 // src-size 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifndef PANEL_MINB
#define PANEL_MINB 2          // CTAs per SM of the streaming (128-row) variant: <= 128 registers
#endif
template <int ROWS>
__global__ void __launch_bounds__(PANEL_THREADS, PANEL_MINB) k_panel(DevCtx cx, const int4* __restrict__ tasks) {
    extern __shared__ __align__(16) double dsm[];
    // DW[0] = W: W[p][c] = U[p][c], the factored block as the kind-0 tiles read it.
    // DW[1] = D: the raw diagonal block D[i][c]; once warp 0 holds it in registers, D[p][c] = L[c][p] for kinds 1, 2.
    __shared__ __align__(16) double DW[2][NB][CLD];
    double (*const W)[CLD] = DW[0];
    double (*const D)[CLD] = DW[1];
    __shared__ double rd[NB];
    TRACEP(0);
    pdl_trigger();
    int4 tk = tasks[blockIdx.x];
    const Front F = load_front(cx, tk.x);
    pdl_wait();
    const int g = tk.y & 15, ngroup = (tk.y >> 4) & 4095, total = (tk.y >> 16) & 0x3fff;
    // mode 0: all three kinds of tiles (one GPU).  Partitioned top front: mode 1 = the panel owner's pass, kinds 0 and 2
    // (the pivot block and L21); mode 2 = every rank's pass over the rows of U12' it owns (kind 1), with the diagonal
    // block already factored in P (replicated by the owner).
    const int mode = (int)((unsigned)tk.y >> 30);
    const int k = F.k, j0 = g * NB, w = (k - j0 < NB) ? k - j0 : NB, j1 = j0 + w;
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id(), fr = lane >> 2, fc = lane & 3;
    // The step's tiles in canonical order: kind 0 (rows below the diagonal block), kind 1 (rows of U12'), kind 2
    // (columns right of the diagonal block).  This CTA takes tiles [tk.z, tk.z + ngroup).
    const int nt0 = (int)((F.f - j1 + ROWS - 1) / ROWS), nt1 = (int)((F.r + ROWS - 1) / ROWS);
    int kind, tile;
    int64_t stride;
    const double* cf;
    auto tile_of = [&](int t, int& kd, int& tl) {
        if (mode == 1) { kd = t < nt0 ? 0 : 2; tl = kd == 0 ? t : t - nt0; }
        else if (mode == 2) { kd = 1; tl = t; }
        else {
            kd = t < nt0 ? 0 : (t < nt0 + nt1 ? 1 : 2);
            if (nt0 + nt1 + (k - j1 + ROWS - 1) / ROWS == 0) kd = 0;        // the lone CTA that only factors D_gg
            tl = kd == 0 ? t : (kd == 1 ? t - nt0 : t - nt0 - nt1);
        }
    };
    auto set_tile = [&](int t) {
        tile_of(t, kind, tile);
        stride = kind == 0 ? F.f : (kind == 1 ? F.r : 1);
    };
    // pull the rows of tile t (columns [0, j1) of the row strip: ROWS * 8 bytes = 8 or 9 lines per column) towards L2
    // while the CTA works on the tile before it
    auto prefetch_tile = [&](int t) {
        if (ROWS == PANEL_ROWS_TOP || t >= (int)tk.z + ngroup) return;
        int kd, tl;
        tile_of(t, kd, tl);
        if (kd == 2) return;
        const int64_t r0 = (int64_t)tl * ROWS, ld = kd == 0 ? F.f : F.r;
        const int64_t nrows = (kd == 0 ? F.f - j1 : F.r) - r0;
        const double* __restrict__ base = (kd == 0 ? F.P + j1 : F.T) + r0;
        for (int l = threadIdx.x; l < j1 * 9; l += PANEL_THREADS) {
            const int col = l / 9;
            const int64_t ro = (l - col * 9) * 16;
            if (ro < nrows && ro < ROWS + 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)col * ld + ro));
        }
    };
    set_tile(tk.z);
    prefetch_tile((int)tk.z + 1);
    TRACEP(1);
    double* Uc = dsm;                               // Uc[m * CLD + c] = U[m, j0 + c],  m < j0
    double* Lc = dsm + j0 * CLD;                    // Lc[m * CLD + i] = L[j0 + i, m],  m < j0
    double* Xs = dsm + 2 * j0 * CLD;                // Xs[row][CLD]: updated rows, one per thread in phase T
    // Latency variant (ROWS == 32): the CTA's 32 rows, columns [0, j1), are staged in shared memory as well
    // (As[row][ALD], row stride = 4 mod 16: conflict-free A fragments), so that the update phase never waits
    // on global memory.
    constexpr bool STAGE_ROWS = ROWS == PANEL_ROWS_TOP;
    constexpr int ALD = KW + 4;
    double* As = Xs + ROWS * CLD;
    const signed char* __restrict__ rown = cx.rowown + cx.rows_ptr[tk.x];
    auto row_ptr = [&](int64_t idx, double*& b, bool& act) {
        if (kind == 0) { act = j1 + idx < F.f; b = F.P + j1 + idx; }
        else if (kind == 1) { act = idx < F.r && (mode != 2 || rown[idx] == cx.rank); b = F.T + idx; }
        else { act = j1 + idx < k; b = F.P + (j1 + idx) * F.f; }
    };
    // ---- S: stage D_gg and the coefficient blocks
    {
        const double* __restrict__ Pg = F.P;
        constexpr int NW = PANEL_THREADS / 32;
        if (STAGE_ROWS) {
            double* rb; bool ra;
            row_ptr((int64_t)tile * ROWS + lane, rb, ra);
            for (int m = warp; m < j1; m += NW) cp_async8(As + lane * ALD + m, ra ? rb + (int64_t)m * stride : Pg, ra);
        }
        {
            const bool ok = lane < w;
            for (int c = warp; c < NB; c += NW)                           // D[i][c], lanes = rows
                cp_async8(&D[lane][c], Pg + ((ok && c < w) ? (j0 + lane) + (int64_t)(j0 + c) * F.f : 0), ok && c < w);
            for (int m = warp; m < j0; m += NW)                           // Lc[m][i] = L[j0+i, m], lanes = rows
                cp_async8(Lc + m * CLD + lane, Pg + (ok ? (j0 + lane) + (int64_t)m * F.f : 0), ok);
        }
        for (int c = warp; c < NB; c += NW) {                             // Uc[m][c] = U[m, j0+c], lanes = m
            const bool ok = c < w;
            const double* __restrict__ src = Pg + (ok ? (int64_t)(j0 + c) * F.f : 0);
            for (int m = lane; m < j0; m += 32) cp_async8(Uc + m * CLD + c, src + (ok ? m : 0), ok);
        }
        cp_async_wait_all();
    }
    __syncthreads();
    TRACEP(2);
    // ---- U: left-looking update on the FP64 tensor pipe, in pieces of 8 rows x 16 columns (2 DMMA tiles).
    // The diagonal block goes first (one piece per warp); then warp 0 factors it (phase L) WHILE the other seven
    // warps update the row block, which does not depend on the factorization.
    cf = kind == 0 ? Uc : Lc;
    constexpr int NW = PANEL_THREADS / 32, ROW_PIECES = STAGE_ROWS ? (ROWS / 8) * 2 : ROWS / 8;   // 8 x 16 (staged) or 8 x 32 pieces
    auto row_piece = [&](int st, int ch) {
        double acc[2][2];
        if (STAGE_ROWS) {
            const double* __restrict__ ar = As + (st * 8 + fr) * ALD;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = ch + 8 * j + 2 * fc + e;
                    acc[j][e] = c < w ? ar[j0 + c] : 0.0;
                }
            for (int m0 = 0; m0 < j0; m0 += 4) {
                const double a = -ar[m0 + fc];
                const double* __restrict__ cm = cf + (m0 + fc) * CLD + ch + fr;
                dmma884(acc[0][0], acc[0][1], a, cm[0]);
                dmma884(acc[1][0], acc[1][1], a, cm[8]);
            }
        } else {
        // Streaming variant: one piece = 8 rows x all 32 columns (four DMMA tiles share every A fragment).  The A
        // fragments come straight from global memory in batches of eight k-steps (32 columns of the row strip), the
        // next batch requested before the current one is consumed: 8-16 loads in flight per lane instead of 2.
        double* fb; bool fa;
        row_ptr((int64_t)tile * ROWS + st * 8 + fr, fb, fa);
        double ac[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * j + 2 * fc + e;
                ac[j][e] = (fa && c < w) ? fb[(int64_t)(j0 + c) * stride] : 0.0;
            }
        if (j0 > 0) {
            double av[8], aw[8];
            const double* __restrict__ fp = fb + (int64_t)fc * stride;
            const int64_t st4 = 4 * stride;
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = fa ? fp[u * st4] : 0.0;
            for (int mb = 0; mb < j0; mb += NB) {
                fp += 8 * st4;
                if (mb + NB < j0) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) aw[u] = fa ? fp[u * st4] : 0.0;
                }
                const double* __restrict__ cm = cf + (mb + fc) * CLD + fr;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double a = -av[u];
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(ac[j][0], ac[j][1], a, cm[u * 4 * CLD + 8 * j]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) av[u] = aw[u];
            }
        }
        double* xs = Xs + (st * 8 + fr) * CLD + 2 * fc;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(xs + 8 * j) = make_double2(ac[j][0], ac[j][1]);
        return;
        }
        double* xs = Xs + (st * 8 + fr) * CLD + ch + 2 * fc;
#pragma unroll
        for (int j = 0; j < 2; ++j) *reinterpret_cast<double2*>(xs + 8 * j) = make_double2(acc[j][0], acc[j][1]);
    };
    auto diag_piece = [&](int i0, int ch) {      // D -= L[g, 0:j0] U[0:j0, g]
        double acc[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(&D[i0 + fr][ch + 8 * j + 2 * fc]);
            acc[j][0] = v.x; acc[j][1] = v.y;
        }
        for (int m0 = 0; m0 < j0; m0 += 4) {
            const double a = -Lc[(m0 + fc) * CLD + i0 + fr];
            const double* __restrict__ cm = Uc + (m0 + fc) * CLD + ch + fr;
            dmma884(acc[0][0], acc[0][1], a, cm[0]);
            dmma884(acc[1][0], acc[1][1], a, cm[8]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(&D[i0 + fr][ch + 8 * j + 2 * fc]) = make_double2(acc[j][0], acc[j][1]);
    };
    if (j0 > 0 && mode != 2) diag_piece((warp >> 1) * 8, (warp & 1) * 16);       // 8 pieces, 8 warps
    static_assert(NW == 8, "one diagonal piece per warp");
    __syncthreads();
    TRACEP(3);
    double x[NB];
    // ---- L: warp 0 factors the diagonal block
    if (warp == 0) {
#pragma unroll
        for (int c2 = 0; c2 < NB / 2; ++c2) {
            const double2 v = *reinterpret_cast<const double2*>(&D[lane][2 * c2]);
            x[2 * c2] = v.x; x[2 * c2 + 1] = v.y;
        }
        // a partial block (w < 32) is padded with the identity, so that all 32 steps run without a special case
#pragma unroll
        for (int c = 0; c < NB; ++c) x[c] = (lane >= w && c == lane) ? 1.0 : x[c];
        double myr = 0.0;                          // 1 / u_jj of this lane's row
        int bad = NB;                              // first bad pivot (uniform across the warp)
        bool big_l = false;                        // threshold test: some multiplier above 1 / pivot_tol
        if (mode != 2) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const double piv = __shfl_sync(0xffffffffu, x[j], j);
            const double rinv = 1.0 / piv;
            bad = (bad == NB && bad_pivot(piv)) ? j : bad;
            myr = lane == j ? rinv : myr;
            const double l = lane > j ? x[j] * rinv : 0.0;
            big_l |= fabs(l) > cx.lmax;
            x[j] = lane > j ? l : x[j];
#pragma unroll
            for (int c = j + 1; c < NB; ++c) x[c] -= l * __shfl_sync(0xffffffffu, x[c], j);
        }
        }
        rd[lane] = myr;
        if (big_l && lane < w) atomicMin(cx.flag + 1, F.c0 + j0);
        if (lane == 0 && bad < w) atomicMin(cx.flag, F.c0 + j0 + bad);
        // publish the factors in the form(s) this CTA's tiles need: rows of U for kind 0 (in W), columns of L for
        // kinds 1 and 2 (in D, whose raw contents only this warp still needed)
        const bool need_u = kind == 0, need_l = kind != 0 || (int)tk.z + ngroup > nt0;
        __syncwarp();
        if (need_l) {                              // D[p][c] = L[c][p]: lane c stores column c
#pragma unroll
            for (int p = 0; p < NB; ++p) D[p][lane] = x[p];
        }
        if (need_u) {                              // W[p][c] = U[p][c]: lane p stores its row
#pragma unroll
            for (int c2 = 0; c2 < NB / 2; ++c2) *reinterpret_cast<double2*>(&W[lane][2 * c2]) = make_double2(x[2 * c2], x[2 * c2 + 1]);
        }
    } else {
        for (int pc = warp - 1; pc < ROW_PIECES; pc += NW - 1) row_piece(STAGE_ROWS ? pc >> 1 : pc, STAGE_ROWS ? (pc & 1) * 16 : 0);
    }
    TRACEP(4);
    __syncthreads();
    TRACEP(5);
    if (warp == 0 && mode != 2) {
        // this CTA's reads of the raw diagonal block completed in phase S; the CTA that arrives last
        // stores the factors (column by column: lanes = rows) and the reciprocal pivots
        int last = 0;
        if (lane == 0) {
            __threadfence();
            const int old = atomicAdd(cx.counters + tk.w, 1);
            last = ((old + 1) % total) == 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && lane < w) {
            double* __restrict__ dst = F.P + (j0 + lane) + (int64_t)j0 * F.f;
#pragma unroll
            for (int c = 0; c < NB; ++c) if (c < w) dst[(int64_t)c * F.f] = x[c];
            cx.dinv[F.c0 + j0 + lane] = rd[lane];
        }
    }
    // ---- T: thread = row; warp 0 is busy storing the block, so the rows start at warp 1 when ROWS < 224
    constexpr int T0 = ROWS <= PANEL_THREADS - 32 ? 32 : 0;
    const int row = tid - T0;
    for (int t = tk.z;;) {
        if (row >= 0 && row < ROWS) {
            double* base; bool active;
            row_ptr((int64_t)tile * ROWS + row, base, active);
            if (active) {
#pragma unroll
                for (int c2 = 0; c2 < NB / 2; ++c2) {
                    const double2 v = *reinterpret_cast<const double2*>(Xs + row * CLD + 2 * c2);
                    x[2 * c2] = v.x; x[2 * c2 + 1] = v.y;
                }
                const double (*Wk)[CLD] = DW[kind == 0 ? 0 : 1];
                bool big_l = false;
#pragma unroll
                for (int p = 0; p < NB; ++p) {
                    if (p >= w) break;
                    const double xp = kind == 0 ? x[p] * rd[p] : x[p];
                    big_l |= kind == 0 && fabs(xp) > cx.lmax;      // rows of L: multipliers
                    x[p] = xp;
                    const double2* __restrict__ wr = reinterpret_cast<const double2*>(&Wk[p][0]);
#pragma unroll
                    for (int c2 = (p + 1) / 2; c2 < NB / 2; ++c2) {
                        const double2 wv = wr[c2];
                        if (2 * c2 > p) x[2 * c2] -= xp * wv.x;
                        x[2 * c2 + 1] -= xp * wv.y;
                    }
                }
#pragma unroll
                for (int c = 0; c < NB; ++c) if (c < w) base[(int64_t)(j0 + c) * stride] = x[c];
                if (big_l) atomicMin(cx.flag + 1, F.c0 + j0);
            }
        }
        TRACEP(6);
        // ---- further tiles of this CTA's group (bulk levels): same coefficient blocks, same factored D_gg
        if (STAGE_ROWS || ++t >= (int)tk.z + ngroup) break;    // the latency variant has one tile per CTA
        __syncthreads();                           // phase T has consumed Xs
        set_tile(t);
        prefetch_tile(t + 1);
        cf = kind == 0 ? Uc : Lc;
        for (int pc = warp; pc < ROW_PIECES; pc += NW) row_piece(pc, 0);      // (only the streaming variant has groups)
        __syncthreads();
    }
}

// ------------------------------------------------------------------ Schur update of the CB
// V (r x r) = beta*C - L21 (r x k) * U12 (k x r), U12 held transposed, k <= KW.  64x64 tile per
// CTA; each of the 8 warps owns 32 x 16 of it as 4 x 2 DMMA (m8n8k4) accumulator tiles; K is staged
// through shared memory in chunks of 32 (L21 negated on the way in, so the MMA accumulates C - L21 U12).
// task: x = supernode, y = tile row, z = tile col, w = flags:
//   bit0 beta   : C holds assembled contributions (else it is taken as zero)
//   bit1 direct : V is written straight into the parent's front through the rel map (this front
//                 is its parent's only child, so nobody else writes there): panel entries +=,
//   bit2 assign : ... and the parent's contribution-block entries are assigned (=) instead of +=.
//   without bit1 V overwrites C in place.
constexpr int GEMM_LDS = GEMM_TILE + 4;   // row stride = 4 (mod 16) doubles: fragment loads are conflict-free

#ifndef GEMM_MINB
#define GEMM_MINB 4             // CTAs per SM (<= 64 registers): measured 2 -> 944 ms, 3 -> 886 ms, 4 (16-column chunks) -> 822 ms of Schur updates at 128^3
#endif
#ifndef GEMM_KC
#define GEMM_KC 16              // pivot columns per staged operand chunk of k_gemm_cb (2 stages = 34 KB per CTA: four CTAs per SM)
#endif
constexpr int GKC = GEMM_KC;
#ifndef GEMM_STAGES
#define GEMM_STAGES 2           // cp.async stages of the operand chunks (3: one barrier per chunk instead of two)
#endif
#ifndef GEMM_AHEAD
#define GEMM_AHEAD 592          // tiles ahead whose contribution-block tile is pulled into L2 (0 = off): one wave of CTAs (4 per SM)
#endif
__global__ void __launch_bounds__(256, GEMM_MINB) k_gemm_cb(DevCtx cx, const int4* __restrict__ tasks, int ntasks) {
    extern __shared__ __align__(16) double gsm[];         // 2 stages x (As[GKC][GEMM_LDS] | Bs[GKC][GEMM_LDS])
    pdl_trigger();
    int4 tk = tasks[blockIdx.x];
#if GEMM_AHEAD > 0
    // the task two waves ahead (its C tile is pulled towards L2 below): requested together with this CTA's own task, so the
    // round trip is not on the tile's critical path
    int4 t2 = make_int4(-1, 0, 0, 0);
    if ((int)blockIdx.x + GEMM_AHEAD < ntasks) t2 = tasks[blockIdx.x + GEMM_AHEAD];
#endif
    const Front F = load_front(cx, tk.x);
    pdl_wait();
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile columns: tk.z = tile index; or (flag bit6, distributed top) tk.z = first column and bits 8..14 of the flags =
    // width (<= 64): the tiling restarts wherever the column owner changes, so no tile is computed by two ranks
    const bool colrun = tk.w & 64;
    const int64_t m0 = (int64_t)tk.y * GEMM_TILE, n0 = colrun ? (int64_t)tk.z : (int64_t)tk.z * GEMM_TILE;
    const int64_t ncend = colrun ? n0 + ((tk.w >> 8) & 127) : F.r;            // columns [n0, ncend) exist (ncend <= F.r)
    const double* __restrict__ A = F.P + F.k;
    const double* __restrict__ B = F.T;
    const bool beta = tk.w & 1, direct = tk.w & 2, assign = tk.w & 4;
#ifdef GEMM_EARLY_PARENT
    // the parent's geometry (direct epilogue) is requested now: two dependent round trips less at the end of the tile
    Front Q = F;
    if (direct) Q = load_front(cx, cx.sn_parent[tk.x]);
#endif
    // partitioned factorization: bit3 = this front is a subtree root below the cut: column b of V goes to the rank that
    // owns global column rows[b], into that rank's copy of the block (peer memory over NVLink, or this rank's own);
    // bit4 = this front is in the distributed top: only the columns this rank owns are updated
    const bool xchg = tk.w & 8, mine_only = tk.w & 16;
    const signed char* __restrict__ rown = cx.rowown + cx.rows_ptr[tk.x];
    // warp tile: 32 rows x 16 columns = 4 x 2 DMMA tiles; 8 warps cover 64 x 64
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
    const int fr = lane >> 2, fc = lane & 3;            // fragment row / k (A), n / k (B), row / column pair (C)
    // bit5 = the child's row list IS the parent's front (a link of a chain of fronts cut out of one wide separator):
    // entry (row, col) of V lands at front position (row, col) of the parent, no index map needed
    const bool ident = tk.w & 32;
    // interior tile: all 64 x 64 entries exist, no bounds checks in the prologue / epilogue
    const bool full = m0 + GEMM_TILE <= F.r && n0 + GEMM_TILE <= ncend;
    // stage one K chunk (32 columns of L21, 32 rows of U12) with cp.async; rows beyond the block are zero-filled.
    // Thread (a, p0) copies element a of columns p0, p0 + 4, ... of the chunk: running pointers, no index arithmetic.
    const int sa = tid & (GEMM_TILE - 1), sp0 = tid >> 6;
    const bool oka = m0 + sa < F.r, okb = n0 + sa < ncend;
    const double* __restrict__ Ath = A + (oka ? m0 + sa : 0) + (int64_t)sp0 * F.f;
    const double* __restrict__ Bth = B + (okb ? n0 + sa : 0) + (int64_t)sp0 * F.r;
    const int64_t astep = 4 * F.f, bstep = 4 * F.r;
    auto stage = [&](int kc, int buf) {
        double* as = gsm + buf * (2 * GKC * GEMM_LDS) + sp0 * GEMM_LDS + sa;
        double* bs = as + GKC * GEMM_LDS;
        const double* __restrict__ ap = Ath + (int64_t)kc * F.f;
        const double* __restrict__ bp = Bth + (int64_t)kc * F.r;
        if (kc + GKC <= k) {
#pragma unroll
            for (int u = 0; u < GKC / 4; ++u, ap += astep, bp += bstep) {
                cp_async8(as + u * 4 * GEMM_LDS, ap, oka);
                cp_async8(bs + u * 4 * GEMM_LDS, bp, okb);
            }
        } else {
#pragma unroll
            for (int u = 0; u < GKC / 4; ++u, ap += astep, bp += bstep) {
                const bool in = kc + sp0 + 4 * u < k;
                cp_async8(as + u * 4 * GEMM_LDS, in ? ap : A, in && oka);
                cp_async8(bs + u * 4 * GEMM_LDS, in ? bp : B, in && okb);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[4][2][2];
#ifdef GEMM_C_FIRST
    // the C tile comes from HBM, the operand chunks from L2: request C first
    // the C loads are in flight while the operand tiles arrive
    if (beta && full) {
        const double* __restrict__ c0 = F.C + (m0 + wm + fr) + (n0 + wn + 2 * fc) * F.r;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double* __restrict__ cc = c0 + (8 * j + e) * F.r;
                const bool on = !mine_only || rown[n0 + wn + 8 * j + 2 * fc + e] == cx.rank;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][j][e] = on ? cc[8 * i] : 0.0;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t row = m0 + wm + 8 * i + fr, col = n0 + wn + 8 * j + 2 * fc + e;
                    acc[i][j][e] = (beta && row < F.r && col < ncend && (!mine_only || rown[col] == cx.rank)) ? F.C[row + col * F.r] : 0.0;
                }
    }
    stage(0, 0);
#if GEMM_AHEAD > 0
    // The tile that will run on this SM about two CTA generations from now reads its 64 x 64 piece of C from HBM
    // (the trailing matrix is far larger than L2): pull it towards L2 now, so that its prologue sees an L2 hit.
    // One 128-byte line per thread (64 columns x 4 lines); only when that tile belongs to the same front.
    if (beta) {
        if (t2.x == tk.x) {
            const int64_t pm = (int64_t)t2.y * GEMM_TILE + (tid & 3) * 16, pn = ((t2.w & 64) ? (int64_t)t2.z : (int64_t)t2.z * GEMM_TILE) + (tid >> 2);
            if (pm < F.r && pn < F.r) asm volatile("prefetch.global.L2 [%0];" ::"l"(F.C + pm + pn * F.r));
        }
    }
#endif
#else
    stage(0, 0);
#if GEMM_AHEAD > 0
    // The tile that will run on this SM about two CTA generations from now reads its 64 x 64 piece of C from HBM
    // (the trailing matrix is far larger than L2): pull it towards L2 now, so that its prologue sees an L2 hit.
    // One 128-byte line per thread (64 columns x 4 lines); only when that tile belongs to the same front.
    if (beta) {
        if (t2.x == tk.x) {
            const int64_t pm = (int64_t)t2.y * GEMM_TILE + (tid & 3) * 16, pn = ((t2.w & 64) ? (int64_t)t2.z : (int64_t)t2.z * GEMM_TILE) + (tid >> 2);
            if (pm < F.r && pn < F.r) asm volatile("prefetch.global.L2 [%0];" ::"l"(F.C + pm + pn * F.r));
        }
    }
#endif
    // the C loads are in flight while the operand tiles arrive
    if (beta && full) {
        const double* __restrict__ c0 = F.C + (m0 + wm + fr) + (n0 + wn + 2 * fc) * F.r;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double* __restrict__ cc = c0 + (8 * j + e) * F.r;
                const bool on = !mine_only || rown[n0 + wn + 8 * j + 2 * fc + e] == cx.rank;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][j][e] = on ? cc[8 * i] : 0.0;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t row = m0 + wm + 8 * i + fr, col = n0 + wn + 8 * j + 2 * fc + e;
                    acc[i][j][e] = (beta && row < F.r && col < ncend && (!mine_only || rown[col] == cx.rank)) ? F.C[row + col * F.r] : 0.0;
                }
    }
#endif
    int buf = 0;
#if GEMM_STAGES == 3
    // three stages, one barrier per chunk: chunk kc + 2 GKC is requested right after the barrier that also tells every
    // warp is done with chunk kc - GKC (whose stage it overwrites), so a chunk has two chunks of MMA time to arrive
    if (GKC < k) stage(GKC, 1);
    for (int kc = 0; kc < k; kc += GKC, buf = buf == 2 ? 0 : buf + 1) {
        if (kc + GKC < k) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (kc + 2 * GKC < k) stage(kc + 2 * GKC, buf == 0 ? 2 : buf - 1);
        const double* As = gsm + buf * (2 * GKC * GEMM_LDS);
#else
    for (int kc = 0; kc < k; kc += GKC, buf ^= 1) {
        if (kc + GKC < k) {
            stage(kc + GKC, buf ^ 1);                     // next chunk into the other stage
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* As = gsm + buf * (2 * GKC * GEMM_LDS);
#endif
        const double* Bs = As + GKC * GEMM_LDS;
        const int kw = (k - kc < GKC) ? k - kc : GKC;
        const int ksteps = (kw + 3) >> 2;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int p = 4 * ks + fc;
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = -As[p * GEMM_LDS + wm + 8 * i + fr];      // C - L21 U12
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Bs[p * GEMM_LDS + wn + 8 * j + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
#if GEMM_STAGES != 3
        __syncthreads();                                 // the stage is refilled two iterations from now
#endif
    }
    if (!direct) {
        const int64_t coff = F.C - cx.cb;
        if (full && !xchg) {
            double* __restrict__ c0 = F.C + (m0 + wm + fr) + (n0 + wn + 2 * fc) * F.r;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (mine_only && rown[n0 + wn + 8 * j + 2 * fc + e] != cx.rank) continue;
                    double* __restrict__ cc = c0 + (8 * j + e) * F.r;
#pragma unroll
                    for (int i = 0; i < 4; ++i) cc[8 * i] = acc[i][j][e];
                }
            return;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t col = n0 + wn + 8 * j + 2 * fc + e;
                if (col >= ncend) continue;
                double* __restrict__ dst = F.C;
                if (xchg) dst = cx.cb_peer[rown[col]] + coff;            // the column's owner (maybe this rank)
                else if (mine_only && rown[col] != cx.rank) continue;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t row = m0 + wm + 8 * i + fr;
                    if (row < F.r) dst[row + col * F.r] = acc[i][j][e];
                }
            }
        return;
    }
#ifndef GEMM_EARLY_PARENT
    const Front Q = load_front(cx, cx.sn_parent[tk.x]);
#endif
    if (ident && full && m0 >= Q.k && n0 >= Q.k) {
        // link of a chain, tile inside the parent's contribution block: a shifted copy
        double* __restrict__ d0 = Q.C + (m0 - Q.k + wm + fr) + (n0 - Q.k + wn + 2 * fc) * Q.r;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (mine_only && rown[n0 + wn + 8 * j + 2 * fc + e] != cx.rank) continue;
                double* __restrict__ dd = d0 + (8 * j + e) * Q.r;
                if (assign) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) dd[8 * i] = acc[i][j][e];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) dd[8 * i] += acc[i][j][e];
                }
            }
        return;
    }
    const int* __restrict__ rel = cx.rel + cx.rows_ptr[tk.x];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int64_t col = n0 + wn + 8 * j + 2 * fc + e;
            if (col >= ncend || (mine_only && rown[col] != cx.rank)) continue;
            const int64_t pb = ident ? col : rel[col];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t row = m0 + wm + 8 * i + fr;
                if (row >= F.r) continue;
                const int64_t pa = ident ? row : rel[row];
                const double v = acc[i][j][e];
                if (pb < Q.k) Q.P[pa + pb * Q.f] += v;
                else if (pa < Q.k) Q.T[(pb - Q.k) + pa * Q.r] += v;
                else {
                    double* d = Q.C + (pa - Q.k) + (pb - Q.k) * Q.r;
                    *d = assign ? v : *d + v;
                }
            }
        }
}

// Strip variant of k_gemm_cb: one CTA walks `ni` consecutive row tiles of ONE tile column.  The 64 x k piece of U12
// stays in shared memory for the whole strip (half the operand traffic and copy instructions), the front's geometry is
// read once, and the cp.async pipeline of the L21 chunks never drains: chunk 0 of the next tile is in flight under the
// last chunk and the epilogue of the current one.  task: y = first row tile | tiles << 16; everything else as k_gemm_cb.
__global__ void __launch_bounds__(256, 2) k_gemm_strip(DevCtx cx, const int4* __restrict__ tasks) {
    extern __shared__ __align__(16) double gsm[];         // Bs[KW][GEMM_LDS] | 2 stages x As[NB][GEMM_LDS]
    pdl_trigger();
    const int4 tk = tasks[blockIdx.x];
    const Front F = load_front(cx, tk.x);
    const bool beta = tk.w & 1, direct = tk.w & 2, assign = tk.w & 4, xchg = tk.w & 8, mine_only = tk.w & 16, ident = tk.w & 32, colrun = tk.w & 64;
    Front Q = F;
    if (direct) Q = load_front(cx, cx.sn_parent[tk.x]);   // static geometry: requested before the dependency wait
    pdl_wait();
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ni = (tk.y >> 16) > 0 ? (tk.y >> 16) : 1;
    const int64_t m0s = (int64_t)(tk.y & 0xffff) * GEMM_TILE, n0 = colrun ? (int64_t)tk.z : (int64_t)tk.z * GEMM_TILE;
    const int64_t ncend = colrun ? n0 + ((tk.w >> 8) & 127) : F.r;
    const double* __restrict__ A = F.P + F.k;
    const double* __restrict__ B = F.T;
    const signed char* __restrict__ rown = cx.rowown + cx.rows_ptr[tk.x];
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16, fr = lane >> 2, fc = lane & 3;
    const int sa = tid & (GEMM_TILE - 1), sp0 = tid >> 6;
    const bool okb = n0 + sa < ncend;
    double* const Bs_all = gsm;
    double* const As_all = gsm + KW * GEMM_LDS;
    const int nchunk = (k + NB - 1) / NB;
    {   // the strip's piece of U12 (k x 64), once
        const double* __restrict__ bp = B + (okb ? n0 + sa : 0) + (int64_t)sp0 * F.r;
        double* bs = Bs_all + sp0 * GEMM_LDS + sa;
        for (int p = sp0; p < nchunk * NB; p += 4, bp += 4 * F.r, bs += 4 * GEMM_LDS) cp_async8(bs, p < k ? bp : B, p < k && okb);
    }
    const double* __restrict__ Acol = A + (int64_t)sp0 * F.f;
    const int64_t astep = 4 * F.f;
    auto stage = [&](int q) {              // chunk q of the strip: tile q / nchunk, columns [32 (q % nchunk), +32) of L21
        const int ti = q / nchunk, kc = (q - ti * nchunk) * NB;
        const int64_t mrow = m0s + (int64_t)ti * GEMM_TILE + sa;
        const bool oka = mrow < F.r;
        double* as = As_all + (q & 1) * (NB * GEMM_LDS) + sp0 * GEMM_LDS + sa;
        const double* __restrict__ ap = Acol + (oka ? mrow : 0) + (int64_t)kc * F.f;
#pragma unroll
        for (int u = 0; u < NB / 4; ++u, ap += astep) {
            const bool in = kc + sp0 + 4 * u < k;
            cp_async8(as + u * 4 * GEMM_LDS, in ? ap : A, in && oka);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0);
    double acc[4][2][2];
    const int total = ni * nchunk;
    int64_t m0 = m0s;
    bool full = false;
    for (int q = 0; q < total; ++q) {
        const int ti = q / nchunk, c = q - ti * nchunk;
        if (c == 0) {
            m0 = m0s + (int64_t)ti * GEMM_TILE;
            full = m0 + GEMM_TILE <= F.r && n0 + GEMM_TILE <= ncend;
            // the next tile's piece of C towards L2; this tile's piece into the accumulators (in flight while the operands arrive)
            if (beta && ti + 1 < ni) {
                const int64_t pm = m0 + GEMM_TILE + (tid & 3) * 16, pn = n0 + (tid >> 2);
                if (pm < F.r && pn < ncend) asm volatile("prefetch.global.L2 [%0];" ::"l"(F.C + pm + pn * F.r));
            }
            if (beta && full) {
                const double* __restrict__ c0 = F.C + (m0 + wm + fr) + (n0 + wn + 2 * fc) * F.r;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double* __restrict__ cc = c0 + (8 * j + e) * F.r;
                        const bool on = !mine_only || rown[n0 + wn + 8 * j + 2 * fc + e] == cx.rank;
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[i][j][e] = on ? cc[8 * i] : 0.0;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int64_t row = m0 + wm + 8 * i + fr, col = n0 + wn + 8 * j + 2 * fc + e;
                            acc[i][j][e] = (beta && row < F.r && col < ncend && (!mine_only || rown[col] == cx.rank)) ? F.C[row + col * F.r] : 0.0;
                        }
            }
        }
        if (q + 1 < total) {
            stage(q + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        {
            const double* As = As_all + (q & 1) * (NB * GEMM_LDS);
            const double* Bs = Bs_all + c * NB * GEMM_LDS;
            const int kw = (k - c * NB < NB) ? k - c * NB : NB;
            const int ksteps = (kw + 3) >> 2;
            for (int ks = 0; ks < ksteps; ++ks) {
                const int p = 4 * ks + fc;
                double a[4], b[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = -As[p * GEMM_LDS + wm + 8 * i + fr];      // C - L21 U12
#pragma unroll
                for (int j = 0; j < 2; ++j) b[j] = Bs[p * GEMM_LDS + wn + 8 * j + fr];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        __syncthreads();                                 // the stage is refilled by the next iteration's copy
        if (c != nchunk - 1) continue;
        // ---- epilogue of tile ti (as in k_gemm_cb)
        if (!direct) {
            const int64_t coff = F.C - cx.cb;
            if (full && !xchg) {
                double* __restrict__ c0 = F.C + (m0 + wm + fr) + (n0 + wn + 2 * fc) * F.r;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (mine_only && rown[n0 + wn + 8 * j + 2 * fc + e] != cx.rank) continue;
                        double* __restrict__ cc = c0 + (8 * j + e) * F.r;
#pragma unroll
                        for (int i = 0; i < 4; ++i) cc[8 * i] = acc[i][j][e];
                    }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t col = n0 + wn + 8 * j + 2 * fc + e;
                    if (col >= ncend) continue;
                    double* __restrict__ dst = F.C;
                    if (xchg) dst = cx.cb_peer[rown[col]] + coff;
                    else if (mine_only && rown[col] != cx.rank) continue;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t row = m0 + wm + 8 * i + fr;
                        if (row < F.r) dst[row + col * F.r] = acc[i][j][e];
                    }
                }
            continue;
        }
        if (ident && full && m0 >= Q.k && n0 >= Q.k) {
            double* __restrict__ d0 = Q.C + (m0 - Q.k + wm + fr) + (n0 - Q.k + wn + 2 * fc) * Q.r;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (mine_only && rown[n0 + wn + 8 * j + 2 * fc + e] != cx.rank) continue;
                    double* __restrict__ dd = d0 + (8 * j + e) * Q.r;
                    if (assign) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) dd[8 * i] = acc[i][j][e];
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) dd[8 * i] += acc[i][j][e];
                    }
                }
            continue;
        }
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[tk.x];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t col = n0 + wn + 8 * j + 2 * fc + e;
                if (col >= ncend || (mine_only && rown[col] != cx.rank)) continue;
                const int64_t pb = ident ? col : rel[col];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t row = m0 + wm + 8 * i + fr;
                    if (row >= F.r) continue;
                    const int64_t pa = ident ? row : rel[row];
                    const double v = acc[i][j][e];
                    if (pb < Q.k) Q.P[pa + pb * Q.f] += v;
                    else if (pa < Q.k) Q.T[(pb - Q.k) + pa * Q.r] += v;
                    else {
                        double* d = Q.C + (pa - Q.k) + (pb - Q.k) * Q.r;
                        *d = assign ? v : *d + v;
                    }
                }
            }
    }
}

// ------------------------------------------------------------------ inverses of the diagonal blocks
// After the factorization, one warp per 32 x 32 diagonal block D_gg = L_gg U_gg of a big front's pivot block:
// dblk[block] (32 x 32, column-major) holds L_gg^{-1} below the diagonal (its unit diagonal is implicit) and
// U_gg^{-1} on and above it.  The solve kernels then apply a diagonal block as ONE 32-term product per row
// (independent shuffles) instead of a 32-step dependent substitution chain -- the chain was 3.3 k cycles per
// block, four blocks per 128-column front, on the critical path of every solve launch.
// Lane c computes column c of both inverses by column-oriented substitution; the coefficients are read as
// shared-memory broadcasts, all updates of one step are independent.  A partial block is padded with the identity.
// task: x = supernode, y = block g.
constexpr int INV_WARPS = 4;
__global__ void __launch_bounds__(32 * INV_WARPS) k_diag_inverse(DevCtx cx, const int2* __restrict__ tasks, int ntasks) {
    __shared__ double Bs[INV_WARPS][NB][NB + 1];
    __shared__ double rds[INV_WARPS][NB];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int ti = blockIdx.x * INV_WARPS + wl;
    if (ti >= ntasks) return;
    const int2 tk = tasks[ti];
    const Front F = load_front(cx, tk.x);
    const int j0 = tk.y * NB, w = (F.k - j0 < NB) ? F.k - j0 : NB;
    double (*B)[NB + 1] = Bs[wl];
    {
        const double* __restrict__ src = F.P + (j0 + lane) + (int64_t)j0 * F.f;
#pragma unroll 8
        for (int c = 0; c < NB; ++c) B[lane][c] = (lane < w && c < w) ? src[(int64_t)c * F.f] : (lane == c ? 1.0 : 0.0);
        rds[wl][lane] = lane < w ? 1.0 / src[(int64_t)lane * F.f] : 1.0;     // = dinv (same division as in the factor kernels)
    }
    __syncwarp();
    double x[NB], y[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { x[i] = i == lane ? 1.0 : 0.0; y[i] = x[i]; }
#pragma unroll
    for (int j = 0; j < NB - 1; ++j)                 // L x = e_lane
#pragma unroll
        for (int i = j + 1; i < NB; ++i) x[i] -= B[i][j] * x[j];
#pragma unroll
    for (int j = NB - 1; j >= 0; --j) {              // U y = e_lane
        y[j] *= rds[wl][j];
#pragma unroll
        for (int i = 0; i < j; ++i) y[i] -= B[i][j] * y[j];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NB; ++i) B[i][lane] = i > lane ? x[i] : y[i];
    __syncwarp();
    double* __restrict__ dst = cx.dblk + ((int64_t)cx.Doff[tk.x] + tk.y) * (NB * NB);
#pragma unroll 8
    for (int c = 0; c < NB; ++c) dst[lane + NB * c] = B[lane][c];
}

// ------------------------------------------------------------------ solves
// All solve kernels are templated on RB, the number of right-hand sides swept together (1, 4 or 8): the
// work vectors are interleaved (entry i of right-hand side q at [i * RB + q]), so the factor entries are
// read once per RB right-hand sides and every thread's RB values are contiguous.
// nv <= RB columns are real; the sweep's remaining slots are zero-filled on the way in and dropped on the way out
template <int RB>
__global__ void k_permute_scale(int n, const int* __restrict__ p, const double* __restrict__ Rs,
                                const double* __restrict__ b, int64_t ldb, double* __restrict__ w, int nv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pi = p ? p[i] : i;
    const double sc = Rs ? Rs[pi] : 1.0;
#pragma unroll
    for (int q = 0; q < RB; ++q) w[(int64_t)i * RB + q] = q < nv ? sc * b[pi + q * ldb] : 0.0;
}
template <int RB>
__global__ void k_unpermute(int n, const int* __restrict__ qv, const double* __restrict__ w, double* __restrict__ x, int64_t ldx, int nv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int qi = qv ? qv[i] : i;
#pragma unroll
    for (int q = 0; q < RB; ++q) if (q < nv) x[qi + q * ldx] = w[(int64_t)i * RB + q];
}

// Triangular solves with the k x k pivot block (k <= KW = 128) of a big front, without staging the block:
// thread t < 128 owns row t and keeps its current right-hand-side values in registers; the pivot block is
// consumed 32 columns (block g) at a time, every thread loading its 32 coefficients of the block straight
// from global memory into registers one block ahead (each entry of the pivot block is used exactly once).
// Warp g owns the rows of diagonal block g: it runs the 32-step substitution with shuffles only (the
// coefficients that do not apply are loaded as zeros, so the loop has no predicates); the warps on the far
// side of the block then apply the block's solution values, published in shared memory, to their rows.
__device__ __forceinline__ void diag_load_lower(double (&buf)[NB], const Front& F, const double* __restrict__ inv,
                                                int t, int g, int lane, int warp) {
    const int j0 = g * NB;
    if (warp == g) {                                   // the block's own rows: row `lane` of L_gg^{-1}
#pragma unroll
        for (int c = 0; c < NB; ++c) buf[c] = c < lane ? inv[g * (NB * NB) + lane + NB * c] : 0.0;
    } else {
#pragma unroll
        for (int c = 0; c < NB; ++c)
            buf[c] = (t < F.k && j0 + c < F.k && warp > g) ? F.P[t + (int64_t)(j0 + c) * F.f] : 0.0;
    }
}
__device__ __forceinline__ void diag_load_upper(double (&buf)[NB], const Front& F, const double* __restrict__ inv,
                                                int t, int g, int lane, int warp) {
    const int j0 = g * NB;
    if (warp == g) {                                   // row `lane` of U_gg^{-1}
#pragma unroll
        for (int c = 0; c < NB; ++c) buf[c] = c >= lane ? inv[g * (NB * NB) + lane + NB * c] : 0.0;
    } else {
#pragma unroll
        for (int c = 0; c < NB; ++c)
            buf[c] = (t < F.k && j0 + c < F.k && warp < g) ? F.P[t + (int64_t)(j0 + c) * F.f] : 0.0;
    }
}
// y <- L11^{-1} y (unit lower).  ys: shared, KW * RB doubles; the solution is left there.
// inv: the front's inverted diagonal blocks (k_diag_inverse).  Threads t >= KW only take part in the barriers.
template <int RB, bool AHEAD = true>
__device__ __forceinline__ void diag_solve_lower(const Front& F, const double* __restrict__ inv, double (&y)[RB], double* ys, int tid) {
    const int lane = tid & 31, warp = uniform_warp_id(), nblk = (F.k + NB - 1) / NB;
    double cur[NB], nxt[NB];
    if (AHEAD && tid < KW) diag_load_lower(cur, F, inv, tid, 0, lane, warp);
    for (int g = 0; g < nblk; ++g) {
        if constexpr (AHEAD) { if (tid < KW && g + 1 < nblk) diag_load_lower(nxt, F, inv, tid, g + 1, lane, warp); }
        else if (tid < KW) diag_load_lower(cur, F, inv, tid, g, lane, warp);      // throughput variant: half the registers
        if (warp == g) {
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                double a[4] = {y[q], 0.0, 0.0, 0.0};             // y_i + sum_{j<i} (L^{-1})_ij y_j, fixed order
#pragma unroll
                for (int j = 0; j < NB; ++j) a[j & 3] += cur[j] * __shfl_sync(0xffffffffu, y[q], j);
                y[q] = (a[0] + a[1]) + (a[2] + a[3]);
                ys[tid * RB + q] = y[q];
            }
        }
        __syncthreads();
        if (tid < KW && warp > g) {
            const double* __restrict__ yb = ys + g * NB * RB;
            if (RB == 1) {                                   // four short dependent chains instead of one
                double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int c = 0; c < NB; ++c) a[c & 3] += cur[c] * yb[c * RB];
                y[0] -= (a[0] + a[1]) + (a[2] + a[3]);
            } else {                                         // RB independent chains already
#pragma unroll
                for (int c = 0; c < NB; ++c)
#pragma unroll
                    for (int q = 0; q < RB; ++q) y[q] -= cur[c] * yb[c * RB + q];
            }
        }
        if constexpr (AHEAD) {
#pragma unroll
            for (int c = 0; c < NB; ++c) cur[c] = nxt[c];
        }
    }
}
// x <- U11^{-1} v (upper).  xs: shared, KW * RB doubles; the solution is left there.
template <int RB, bool AHEAD = true>
__device__ __forceinline__ void diag_solve_upper(const Front& F, const double* __restrict__ inv, double (&v)[RB], double* xs, int tid) {
    const int lane = tid & 31, warp = uniform_warp_id(), nblk = (F.k + NB - 1) / NB;
    double cur[NB], nxt[NB];
    if (AHEAD && tid < KW) diag_load_upper(cur, F, inv, tid, nblk - 1, lane, warp);
    for (int g = nblk - 1; g >= 0; --g) {
        if constexpr (AHEAD) { if (tid < KW && g > 0) diag_load_upper(nxt, F, inv, tid, g - 1, lane, warp); }
        else if (tid < KW) diag_load_upper(cur, F, inv, tid, g, lane, warp);
        if (warp == g) {
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                double a[4] = {0.0, 0.0, 0.0, 0.0};              // sum_{j>=i} (U^{-1})_ij v_j, fixed order
#pragma unroll
                for (int j = 0; j < NB; ++j) a[j & 3] += cur[j] * __shfl_sync(0xffffffffu, v[q], j);
                v[q] = (a[0] + a[1]) + (a[2] + a[3]);
                xs[tid * RB + q] = v[q];
            }
        }
        __syncthreads();
        if (tid < KW && warp < g) {
            const double* __restrict__ xb = xs + g * NB * RB;
            if (RB == 1) {                                   // four short dependent chains instead of one
                double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int c = 0; c < NB; ++c) a[c & 3] += cur[c] * xb[c * RB];
                v[0] -= (a[0] + a[1]) + (a[2] + a[3]);
            } else {                                         // RB independent chains already
#pragma unroll
                for (int c = 0; c < NB; ++c)
#pragma unroll
                    for (int q = 0; q < RB; ++q) v[q] -= cur[c] * xb[c * RB + q];
            }
        }
        if constexpr (AHEAD) {
#pragma unroll
            for (int c = 0; c < NB; ++c) cur[c] = nxt[c];
        }
    }
}

// Forward substitution for one level.  task: x = supernode, y = row tile of the update vector.
// y_s = L11^{-1} (w[cols] + children contributions); upd_s = children contributions - L21 y_s.
// Every tile recomputes y_s; tile 0 stores it.  Children are gathered one at a time, ascending,
// so the summation order is fixed.  L11 is applied by diag_solve_lower (coefficients straight from
// global memory into registers).  The tile's FWD_ROWS rows of L21 are reduced by 4 threads per row
// (k/4 columns each, combined in a fixed order).
template <int RB, int ROWS>
__global__ void __launch_bounds__(SOLVE_THREADS, (ROWS >= FWD_ROWS_MID && RB == 1) ? 2 : 1) k_fwd(DevCtx cx, const int4* __restrict__ tasks,
                                                       const double* __restrict__ win, double* __restrict__ zout) {
    __shared__ double ys[KW * RB];
    constexpr int NQ = SOLVE_THREADS / ROWS;                 // threads per row in the L21 product (4, 2 or 1)
    constexpr int QPT = 4 / NQ;                              // quarters of the pivot columns per thread
    __shared__ double acc[ROWS * RB];
    __shared__ double red[NQ > 1 ? 4 : 1][NQ > 1 ? ROWS * RB : 1];
    pdl_trigger();
    int4 tk = tasks[blockIdx.x];
    const int s = tk.x;
    const Front F = load_front(cx, s);
    const int k = F.k, tid = threadIdx.x, kp = ((k + NB - 1) / NB) * NB;
    const int64_t lo = (int64_t)tk.y * ROWS;            // first update row of this tile
    const double* __restrict__ inv = cx.dblk + (int64_t)cx.Doff[s] * (NB * NB);
    prefetch_block_l2(inv, NB, kp, NB, tid, SOLVE_THREADS);
    prefetch_block_l2(F.P, k, k, F.f, tid, SOLVE_THREADS);
    prefetch_block_l2(F.P + k + lo, (int)(F.r - lo < ROWS ? F.r - lo : ROWS), k, F.f, tid, SOLVE_THREADS);
    TRACE2(8);
    pdl_wait();
    TRACE2(9);
    // A single child whose row list is exactly this front (a link of a chain cut out of one wide separator, or the
    // virtual child of an interface front): rel is the identity, so the child's update vector is read in place --
    // pivot rows first, then this tile's rows -- instead of every tile scanning all of it.
    const int ch0 = cx.child_ptr[s], nch = cx.child_ptr[s + 1] - ch0;
    const int c_id = nch == 1 ? cx.child_idx[ch0] : -1;
    const bool ident = nch == 1 && cx.rows_ptr[c_id + 1] - cx.rows_ptr[c_id] == F.f;
    if (ident) {
        const double* __restrict__ uc = cx.upd + cx.rows_ptr[c_id] * RB;
        for (int e = tid; e < KW * RB; e += SOLVE_THREADS) ys[e] = e < k * RB ? win[(int64_t)F.c0 * RB + e] + uc[e] : 0.0;
        const int64_t nrow = F.r - lo < ROWS ? F.r - lo : ROWS;
        for (int e = tid; e < ROWS * RB; e += SOLVE_THREADS) acc[e] = e < nrow * RB ? uc[(k + lo) * RB + e] : 0.0;
        __syncthreads();
    } else {
    for (int e = tid; e < KW * RB; e += SOLVE_THREADS) ys[e] = e < k * RB ? win[(int64_t)F.c0 * RB + e] : 0.0;
    for (int e = tid; e < ROWS * RB; e += SOLVE_THREADS) acc[e] = 0.0;
    __syncthreads();
    }
    for (int ci = ch0; ci < (ident ? ch0 : ch0 + nch); ++ci) {
        const int c = cx.child_idx[ci];
        const int64_t rc = cx.rows_ptr[c + 1] - cx.rows_ptr[c];
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ uc = cx.upd + cx.rows_ptr[c] * RB;
        constexpr int NBATCH = RB >= 4 ? 2 : 4;
        for (int64_t a0 = 0; a0 < rc; a0 += NBATCH * SOLVE_THREADS) {
            int64_t ra[NBATCH]; double uv[NBATCH][RB];
#pragma unroll
            for (int u = 0; u < NBATCH; ++u) {
                const int64_t a = a0 + u * SOLVE_THREADS + tid;
                ra[u] = a < rc ? rel[a] : -1;
#pragma unroll
                for (int q = 0; q < RB; ++q) uv[u][q] = a < rc ? uc[a * RB + q] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < NBATCH; ++u) {
                if (ra[u] < 0) continue;
                if (ra[u] < k) {
#pragma unroll
                    for (int q = 0; q < RB; ++q) ys[ra[u] * RB + q] += uv[u][q];
                } else if (ra[u] - k >= lo && ra[u] - k < lo + ROWS) {
#pragma unroll
                    for (int q = 0; q < RB; ++q) acc[(ra[u] - k - lo) * RB + q] += uv[u][q];
                }
            }
        }
        __syncthreads();
    }
    {
        double y[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) y[q] = tid < k ? ys[tid * RB + q] : 0.0;
        __syncthreads();                                     // everybody has read its entries of ys
        TRACE2(10);
        diag_solve_lower<RB, !(ROWS >= FWD_ROWS_MID && RB == 1)>(F, inv, y, ys, tid);
        __syncthreads();
        TRACE2(11);
    }
    if (tk.y == 0) for (int e = tid; e < k * RB; e += SOLVE_THREADS) zout[(int64_t)F.c0 * RB + e] = ys[e];
    {   // rows of L21: the kp columns are summed in four quarters, combined in a fixed order; with ROWS = 64 the
        // quarters of a row belong to four threads, with ROWS = 256 one thread takes them one after the other
        const int rloc = tid & (ROWS - 1), kq = kp / 4;
        const int64_t row = lo + rloc;
        double vq[QPT][RB];
#pragma unroll
        for (int h = 0; h < QPT; ++h) {
            const int qt = (tid / ROWS) * QPT + h;
            double (&v)[RB] = vq[h];
#pragma unroll
            for (int q = 0; q < RB; ++q) v[q] = 0.0;
            if (row < F.r) {
                const double* __restrict__ src = F.P + F.k + row + (int64_t)(qt * kq) * F.f;
                const int jn = (k - qt * kq < kq) ? k - qt * kq : kq;     // may be <= 0 for the last quarters
                int j = 0;
                if (jn == NB) {                                  // k = 128: the thread's 32 entries in flight at once
                    double l[NB];
#pragma unroll
                    for (int u = 0; u < NB; ++u) l[u] = src[(int64_t)u * F.f];
#pragma unroll
                    for (int u = 0; u < NB; ++u)
#pragma unroll
                        for (int q = 0; q < RB; ++q) v[q] += l[u] * ys[(qt * kq + u) * RB + q];
                    j = NB;
                }
                for (; j + 8 <= jn; j += 8) {
                    double l[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) l[u] = src[(int64_t)(j + u) * F.f];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int q = 0; q < RB; ++q) v[q] += l[u] * ys[(qt * kq + j + u) * RB + q];
                }
                for (; j < jn; ++j) {
                    const double l = src[(int64_t)j * F.f];
#pragma unroll
                    for (int q = 0; q < RB; ++q) v[q] += l * ys[(qt * kq + j) * RB + q];
                }
            }
        }
        if (NQ > 1) {
            const int qt = tid / ROWS;
#pragma unroll
            for (int h = 0; h < QPT; ++h)
#pragma unroll
                for (int q = 0; q < RB; ++q) red[qt * QPT + h][rloc * RB + q] = vq[h][q];
            __syncthreads();
            if (qt == 0 && row < F.r) {
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int e = rloc * RB + q;
                    cx.upd[(cx.rows_ptr[s] + row) * RB + q] = acc[e] - (((red[0][e] + red[1][e]) + red[2][e]) + red[3][e]);
                }
            }
        } else if (row < F.r) {
#pragma unroll
            for (int q = 0; q < RB; ++q)
                cx.upd[(cx.rows_ptr[s] + row) * RB + q] =
                    acc[rloc * RB + q] - (((vq[0][q] + vq[QPT > 1 ? 1 : 0][q]) + vq[QPT > 2 ? 2 : 0][q]) + vq[QPT > 3 ? 3 : 0][q]);
        }
        TRACE2(12);
    }
}

// Backward substitution for one level.  task: x = supernode, y = row tile, z = tiles of this
// supernode, w = slot of its partial sums in cx.bpart.   x[cols] = U11^{-1} (x[cols] - U12 x[rows]).
// Each CTA reduces BWD_ROWS rows of U12' against the gathered x (a warp takes 4 columns at a time
// so 32 loads per lane are in flight); with several tiles the partial k-vectors go to scratch and
// the CTA that arrives last adds them in tile order (fixed summation order, nobody waits) and
// finishes the back substitution (diag_solve_upper: coefficients straight from global memory into registers).
template <int RB, bool WIDE>
__global__ void __launch_bounds__(SOLVE_THREADS, (WIDE && RB == 1) ? 2 : 1) k_bwd(DevCtx cx, const int4* __restrict__ tasks, double* __restrict__ x) {
    __shared__ double xs[BWD_ROWS * RB];
    __shared__ double part[KW * RB];
    __shared__ int s_last;
    pdl_trigger();
    int4 tk = tasks[blockIdx.x];
    // tk.y = row tile | column part << 24 | column parts << 28: on levels with few row tiles the k pivot columns of a tile
    // are split over 2 or 4 CTAs (each reads its columns of the tile of U12'; the partial sums are independent)
    const int s = tk.x, ntiles = tk.z, rtile = tk.y & 0xffffff, cpart = (tk.y >> 24) & 15, nsp = (tk.y >> 28) ? (tk.y >> 28) : 1;
    const Front F = load_front(cx, s);
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
    const int kper = (((k + nsp - 1) / nsp) + 7) & ~7, c_lo = cpart * kper, c_hi = c_lo + kper < k ? c_lo + kper : k;
    const int64_t lo = (int64_t)rtile * BWD_ROWS;
    const int cnt = (int)(F.r - lo < BWD_ROWS ? F.r - lo : BWD_ROWS);
    const double* __restrict__ inv = cx.dblk + (int64_t)cx.Doff[s] * (NB * NB);
    prefetch_block_l2(F.T + lo, cnt, k, F.r, tid, SOLVE_THREADS);
    prefetch_block_l2(inv, NB, ((k + NB - 1) / NB) * NB, NB, tid, SOLVE_THREADS);
    prefetch_block_l2(F.P, k, k, F.f, tid, SOLVE_THREADS);
    TRACE2(0);
    pdl_wait();
    TRACE2(1);
    const int* __restrict__ rows = cx.rows + cx.rows_ptr[s] + lo;
    {
        const int64_t xr = tid < cnt ? (int64_t)rows[tid] * RB : 0;     // BWD_ROWS == SOLVE_THREADS
#pragma unroll
        for (int q = 0; q < RB; ++q) xs[tid * RB + q] = tid < cnt ? x[xr + q] : 0.0;
    }
    __syncthreads();
    TRACE2(2);
    // a warp takes BU columns at a time (8 for a single right-hand side: 64 loads per lane in flight; 4 when
    // RB > 1, where the accumulators need the registers, and in the two-CTAs-per-SM variant of the bulk levels)
    constexpr int BU = (RB == 1 && !WIDE) ? 8 : 4;
    for (int i0 = c_lo + warp * BU; i0 < c_hi; i0 += (SOLVE_THREADS / 32) * BU) {
        double v[BU][RB];
#pragma unroll
        for (int u = 0; u < BU; ++u)
#pragma unroll
            for (int q = 0; q < RB; ++q) v[u][q] = 0.0;
        double t[BU][BWD_ROWS / 32];
#pragma unroll
        for (int u = 0; u < BU; ++u) {
            const double* __restrict__ col = F.T + (int64_t)(i0 + u) * F.r + lo;
#pragma unroll
            for (int a = 0; a < BWD_ROWS / 32; ++a) t[u][a] = (i0 + u < c_hi && lane + 32 * a < cnt) ? col[lane + 32 * a] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < BWD_ROWS / 32; ++a)
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const double xv = xs[(lane + 32 * a) * RB + q];
#pragma unroll
                for (int u = 0; u < BU; ++u) v[u][q] += t[u][a] * xv;
            }
#pragma unroll
        for (int u = 0; u < BU; ++u)
#pragma unroll
            for (int q = 0; q < RB; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[u][q] += __shfl_xor_sync(0xffffffffu, v[u][q], o);
                if (lane == 0 && i0 + u < c_hi) part[(i0 + u) * RB + q] = v[u][q];
            }
    }
    __syncthreads();
    TRACE2(3);
    if (ntiles * nsp > 1) {
        double* slot = cx.bpart + (int64_t)tk.w * KW * RB;
        for (int e = c_lo * RB + tid; e < c_hi * RB; e += SOLVE_THREADS) slot[(int64_t)rtile * KW * RB + e] = part[e];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int old = atomicAdd(cx.counters2 + s, 1);
            s_last = ((old + 1) % (ntiles * nsp)) == 0;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        // tile-ordered sum of the partial vectors: eight loads in flight per thread (the chain of dependent L2 round
        // trips was the longest piece of a level with many tiles), four running sums combined in a fixed order
        for (int e = tid; e < k * RB; e += SOLVE_THREADS) {
            double a4[4] = {0.0, 0.0, 0.0, 0.0};
            const double* __restrict__ sp = slot + e;
            int t = 0;
            for (; t + 8 <= ntiles; t += 8) {
                double l[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) l[u] = __ldcg(sp + (int64_t)(t + u) * KW * RB);
#pragma unroll
                for (int u = 0; u < 8; ++u) a4[u & 3] += l[u];
            }
            for (; t < ntiles; ++t) a4[t & 3] += __ldcg(sp + (int64_t)t * KW * RB);
            part[e] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
        }
        __syncthreads();
    }
    {
        double v0[RB];
        const int64_t xc = (int64_t)(F.c0 + tid) * RB;
#pragma unroll
        for (int q = 0; q < RB; ++q) v0[q] = tid < k ? x[xc + q] - part[tid * RB + q] : 0.0;     // right-hand side of U11 x = ...
        __syncthreads();                                     // everybody has read its entries of part
        TRACE2(4);
        diag_solve_upper<RB, !(WIDE && RB == 1)>(F, inv, v0, part, tid);         // the solution is published in part
        __syncthreads();
        TRACE2(5);
    }
    for (int e = tid; e < k * RB; e += SOLVE_THREADS) x[(int64_t)F.c0 * RB + e] = part[e];
}

// ------------------------------------------------------------------ many right-hand sides: 32-wide sweeps on the FP64 tensor pipe
// With MR = 32 right-hand sides swept together the big-front solves are dense contractions (the ridge is at
// ~37 right-hand sides): upd = gathered - L21 * Y and part = U12 * X run as DMMA (m8n8k4) products whose A fragments
// (the factor entries) come straight from global memory -- every factor entry is read once per 32 right-hand sides --
// and whose B fragments (the right-hand sides) sit in shared memory.  The k x k pivot block is applied 32 columns at a
// time: the inverted diagonal blocks of k_diag_inverse as one 32 x 32 x 32 product, the off-diagonal blocks as further
// products.  Vectors are interleaved [i * 32 + q] like the narrower sweeps.  The small fronts run the 8-wide warp-per-front
// kernels on the four slices of the 32 slots (k_small_fwd / k_small_bwd with LDV = 32).
constexpr int MR = 32;             // right-hand sides per sweep
constexpr int MLD = MR + 4;        // shared-memory row stride: 4 (mod 16) doubles, so B fragment loads are conflict-free
constexpr int MROWS = 256;         // rows of L21 / U12' per CTA (eight warps x 32 rows, or 256 contraction rows backward)

// One warp's share of a 32 x 32 (block) = A (32 x 32, element functor) * B (32 x 32 rows of a shared [..][MLD] array):
// warp w computes tile row i = w >> 1 and the tile columns jb, jb + 1 (jb = 2 (w & 1)); k-steps [ks0, ks1).
template <class FA>
__device__ __forceinline__ void block32_mma(double (&acc)[2][2], FA getA, const double* __restrict__ Bs, int i, int jb,
                                            int fr, int fc, int ks0, int ks1) {
    double a[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) a[ks] = (ks >= ks0 && ks < ks1) ? getA(8 * i + fr, 4 * ks + fc) : 0.0;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        if (ks < ks0 || ks >= ks1) continue;
        const double* __restrict__ bp = Bs + (4 * ks + fc) * MLD + fr;
        dmma884(acc[0][0], acc[0][1], a[ks], bp[8 * jb]);
        dmma884(acc[1][0], acc[1][1], a[ks], bp[8 * jb + 8]);
    }
}
__device__ __forceinline__ void block32_load(double (&acc)[2][2], const double* __restrict__ Cs, int i, int jb, int fr, int fc) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(Cs + (8 * i + fr) * MLD + 8 * (jb + j) + 2 * fc);
        acc[j][0] = v.x; acc[j][1] = v.y;
    }
}
__device__ __forceinline__ void block32_store(const double (&acc)[2][2], double* __restrict__ Cs, int i, int jb, int fr, int fc) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
        *reinterpret_cast<double2*>(Cs + (8 * i + fr) * MLD + 8 * (jb + j) + 2 * fc) = make_double2(acc[j][0], acc[j][1]);
}
// ys (KW x MLD, rows >= k zero) <- L11^{-1} ys (unit lower) or U11^{-1} ys (upper); all 256 threads, ends with a barrier.
template <bool LOWER>
__device__ __forceinline__ void pivot_solve32(const Front& F, const double* __restrict__ inv, double* ys) {
    const int lane = threadIdx.x & 31, w = uniform_warp_id(), fr = lane >> 2, fc = lane & 3;
    const int i = w >> 1, jb = 2 * (w & 1), k = F.k, nblk = (k + NB - 1) / NB;
    const double* __restrict__ P = F.P;
    const int64_t f = F.f;
    for (int gg = 0; gg < nblk; ++gg) {
        const int g = LOWER ? gg : nblk - 1 - gg;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const double* __restrict__ ib = inv + (int64_t)g * (NB * NB);
        if (LOWER) block32_mma(acc, [&](int m, int c) { return m > c ? ib[m + NB * c] : (m == c ? 1.0 : 0.0); }, ys + g * NB * MLD, i, jb, fr, fc, 0, 2 * (i + 1));
        else block32_mma(acc, [&](int m, int c) { return m <= c ? ib[m + NB * c] : 0.0; }, ys + g * NB * MLD, i, jb, fr, fc, 2 * i, 8);
        __syncthreads();                                     // every warp has read the block's right-hand sides
        block32_store(acc, ys + g * NB * MLD, i, jb, fr, fc);
        __syncthreads();
        // the other blocks on the far side: V_h -= L_hg Y_g (h > g) / V_h -= U_hg X_g (h < g)
        for (int h = LOWER ? g + 1 : 0; LOWER ? h < nblk : h < g; ++h) {
            double c2[2][2];
            block32_load(c2, ys + h * NB * MLD, i, jb, fr, fc);
            block32_mma(c2, [&](int m, int c) {
                            const int row = h * NB + m, col = g * NB + c;
                            return (row < k && col < k) ? -P[row + (int64_t)col * f] : 0.0;
                        }, ys + g * NB * MLD, i, jb, fr, fc, 0, 8);
            block32_store(c2, ys + h * NB * MLD, i, jb, fr, fc);
        }
        __syncthreads();
    }
}

// Forward substitution of one level, 32 right-hand sides.  task: x = supernode, y = row tile (TR = 8 WR rows of L21:
// 256 where a level has many tiles, 64 near the top of the tree, where the few fronts of a level are spread over more SMs).
// dynamic shared memory: ys[KW][MLD] | accs[TR][MLD].
template <int WR>
__global__ void __launch_bounds__(SOLVE_THREADS, 2) k_fwd32(DevCtx cx, const int4* __restrict__ tasks,
                                                            const double* __restrict__ win, double* __restrict__ zout) {
    constexpr int TR = 8 * WR, NI = WR / 8;
    extern __shared__ __align__(16) double msm[];
    double* ys = msm;
    double* accs = msm + KW * MLD;
    __shared__ int s_piv[64], s_lo[64], s_hi[64];
    pdl_trigger();
    const int4 tk = tasks[blockIdx.x];
    const int s = tk.x;
    const Front F = load_front(cx, s);
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, w = uniform_warp_id(), fr = lane >> 2, fc = lane & 3;
    const int64_t lo = (int64_t)tk.y * TR;
    const int nrow = (int)(F.r - lo < TR ? (F.r - lo > 0 ? F.r - lo : 0) : TR);
    const double* __restrict__ inv = cx.dblk + (int64_t)cx.Doff[s] * (NB * NB);
    // the factors are constant during a solve: pull this CTA's share towards L2 before waiting for the level below
    prefetch_block_l2(inv, NB, ((k + NB - 1) / NB) * NB, NB, tid, SOLVE_THREADS);
    prefetch_block_l2(F.P, k, k, F.f, tid, SOLVE_THREADS);
    prefetch_block_l2(F.P + k + lo, nrow, k, F.f, tid, SOLVE_THREADS);
    pdl_wait();
    const int ch0 = cx.child_ptr[s], nch = cx.child_ptr[s + 1] - ch0;
    const int c_id = nch == 1 ? cx.child_idx[ch0] : -1;
    const bool ident = nch == 1 && cx.rows_ptr[c_id + 1] - cx.rows_ptr[c_id] == F.f;
    // ---- gather: ys = w[cols] (+ children), accs = children's contributions to this tile's rows; a warp per row
    if (ident) {
        const double* __restrict__ uc = cx.upd + cx.rows_ptr[c_id] * MR;
#pragma unroll 8
        for (int a = w; a < KW; a += 8) ys[a * MLD + lane] = a < k ? win[(int64_t)(F.c0 + a) * MR + lane] + uc[(int64_t)a * MR + lane] : 0.0;
#pragma unroll 8
        for (int a = w; a < TR; a += 8) accs[a * MLD + lane] = a < nrow ? uc[(int64_t)(k + lo + a) * MR + lane] : 0.0;
        __syncthreads();
    } else {
#pragma unroll 8
        for (int a = w; a < KW; a += 8) ys[a * MLD + lane] = a < k ? win[(int64_t)(F.c0 + a) * MR + lane] : 0.0;
        for (int a = w; a < TR; a += 8) accs[a * MLD + lane] = 0.0;
        for (int q0 = 0; q0 < nch; q0 += 64) {
            // rel is ascending: rows [0, piv) of child q land in the pivot rows, rows [lo_q, hi_q) in this tile
            if (tid < 192) {
                const int q = q0 + (tid & 63), what = tid >> 6;
                if (q < nch) {
                    const int c = cx.child_idx[ch0 + q];
                    const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
                    const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
                    const int64_t key = what == 0 ? k : (what == 1 ? k + lo : k + lo + TR);
                    int a0 = 0, a1 = rc;
                    while (a0 < a1) { const int mid = (a0 + a1) >> 1; if (rel[mid] < key) a0 = mid + 1; else a1 = mid; }
                    (what == 0 ? s_piv : (what == 1 ? s_lo : s_hi))[tid & 63] = a0;
                }
            }
            __syncthreads();
            const int q1 = q0 + 64 < nch ? q0 + 64 : nch;
            for (int q = q0; q < q1; ++q) {
                const int c = cx.child_idx[ch0 + q];
                const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
                const double* __restrict__ uc = cx.upd + cx.rows_ptr[c] * MR;
                const int piv = s_piv[q - q0], tlo = s_lo[q - q0], thi = s_hi[q - q0];
                // a warp takes 32 consecutive rows of the child at a time: their positions in one coalesced load, handed
                // round by shuffle; the 32 row loads are independent (eight in flight)
                for (int a0 = 32 * w; a0 < piv; a0 += 256) {
                    const int rl = a0 + lane < piv ? rel[a0 + lane] : 0;
                    const int cntr = piv - a0 < 32 ? piv - a0 : 32;
                    for (int j0 = 0; j0 < cntr; j0 += 8) {
                        double v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = j0 + u < cntr ? uc[(int64_t)(a0 + j0 + u) * MR + lane] : 0.0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int ra = __shfl_sync(0xffffffffu, rl, (j0 + u) & 31);
                            if (j0 + u < cntr) ys[ra * MLD + lane] += v[u];
                        }
                    }
                }
                for (int a0 = tlo + 32 * w; a0 < thi; a0 += 256) {
                    const int rl = a0 + lane < thi ? rel[a0 + lane] - k - (int)lo : 0;
                    const int cntr = thi - a0 < 32 ? thi - a0 : 32;
                    for (int j0 = 0; j0 < cntr; j0 += 8) {
                        double v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = j0 + u < cntr ? uc[(int64_t)(a0 + j0 + u) * MR + lane] : 0.0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int ra = __shfl_sync(0xffffffffu, rl, (j0 + u) & 31);
                            if (j0 + u < cntr) accs[ra * MLD + lane] += v[u];
                        }
                    }
                }
                __syncthreads();                               // children are added one after the other: fixed order
            }
        }
        if (nch == 0) __syncthreads();
    }
    // ---- pivot block
    pivot_solve32<true>(F, inv, ys);
    if (tk.y == 0)
        for (int a = w; a < k; a += 8) zout[(int64_t)(F.c0 + a) * MR + lane] = ys[a * MLD + lane];
    if (nrow == 0) return;
    // ---- upd = accs - L21 Y: warp w owns rows [WR w, WR w + WR) of the tile, NI x 4 DMMA tiles
    double acc[NI][4][2];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(accs + (WR * w + 8 * i + fr) * MLD + 8 * j + 2 * fc);
            acc[i][j][0] = v.x; acc[i][j][1] = v.y;
        }
    if (WR * w < nrow) {
        const double* __restrict__ Lp = F.P + k + lo + WR * w + fr;
        bool ok[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) ok[i] = WR * w + 8 * i + fr < nrow;
        const int ksteps = (k + 3) >> 2;
        for (int ks0 = 0; ks0 < ksteps; ks0 += 4) {
            double a[4][NI];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int col = 4 * (ks0 + u) + fc;
#pragma unroll
                for (int i = 0; i < NI; ++i) a[u][i] = (ok[i] && col < k) ? -Lp[8 * i + (int64_t)col * F.f] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (ks0 + u >= ksteps) continue;
                const double* __restrict__ bp = ys + (4 * (ks0 + u) + fc) * MLD + fr;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double b = bp[8 * j];
#pragma unroll
                    for (int i = 0; i < NI; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[u][i], b);
                }
            }
        }
        double* __restrict__ up = cx.upd + (cx.rows_ptr[s] + lo + WR * w + fr) * MR + 2 * fc;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            if (!ok[i]) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<double2*>(up + (int64_t)(8 * i) * MR + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

// Backward substitution of one level, 32 right-hand sides.  task: x = supernode, y = row tile of U12' (MROWS rows),
// z = tiles of this supernode, w = slot of its partial sums in cx.bpart (KW * MR doubles per tile).
// dynamic shared memory: xs[MROWS][MLD] (reused as the pivot block's right-hand sides) | ps[KW][MLD].
__global__ void __launch_bounds__(SOLVE_THREADS, 2) k_bwd32(DevCtx cx, const int4* __restrict__ tasks, double* __restrict__ x) {
    extern __shared__ __align__(16) double msm[];
    double* xs = msm;
    double* ps = msm + MROWS * MLD;
    __shared__ int s_last;
    pdl_trigger();
    const int4 tk = tasks[blockIdx.x];
    const int s = tk.x, rtile = tk.y, ntiles = tk.z;
    const Front F = load_front(cx, s);
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, w = uniform_warp_id(), fr = lane >> 2, fc = lane & 3;
    const int64_t lo = (int64_t)rtile * MROWS;
    const int cnt = (int)(F.r - lo < MROWS ? (F.r - lo > 0 ? F.r - lo : 0) : MROWS);
    const double* __restrict__ inv = cx.dblk + (int64_t)cx.Doff[s] * (NB * NB);
    const int* __restrict__ rows = cx.rows + cx.rows_ptr[s] + lo;
    prefetch_block_l2(F.T + lo, cnt, k, F.r, tid, SOLVE_THREADS);
    prefetch_block_l2(inv, NB, ((k + NB - 1) / NB) * NB, NB, tid, SOLVE_THREADS);
    prefetch_block_l2(F.P, k, k, F.f, tid, SOLVE_THREADS);
    const int rl = 32 * w + lane < cnt ? rows[32 * w + lane] : 0;      // warp w gathers rows [32 w, 32 w + 32) of the tile
    pdl_wait();
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
        const int ri = __shfl_sync(0xffffffffu, rl, j);
        xs[(32 * w + j) * MLD + lane] = 32 * w + j < cnt ? x[(int64_t)ri * MR + lane] : 0.0;
    }
    __syncthreads();
    // part (k x 32) = U12 (k x rows) X (rows x 32): warp w owns pivot columns [16 w, 16 w + 16), 2 x 4 DMMA tiles
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    if (16 * w < k && cnt > 0) {
        const double* __restrict__ Tp = F.T + lo + fc;
        const bool ok0 = 16 * w + fr < k, ok1 = 16 * w + 8 + fr < k;
        const int64_t c0o = (int64_t)(16 * w + fr) * F.r, c1o = (int64_t)(16 * w + 8 + fr) * F.r;
        const int ksteps = (cnt + 3) >> 2;
        for (int ks0 = 0; ks0 < ksteps; ks0 += 8) {
            double a[8][2];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = 4 * (ks0 + u) + fc;
                a[u][0] = (ok0 && rr < cnt) ? Tp[4 * (ks0 + u) + c0o] : 0.0;
                a[u][1] = (ok1 && rr < cnt) ? Tp[4 * (ks0 + u) + c1o] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (ks0 + u >= ksteps) continue;
                const double* __restrict__ bp = xs + (4 * (ks0 + u) + fc) * MLD + fr;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double b = bp[8 * j];
                    dmma884(acc[0][j][0], acc[0][j][1], a[u][0], b);
                    dmma884(acc[1][j][0], acc[1][j][1], a[u][1], b);
                }
            }
        }
    }
    // partial products -> ps (one tile) or the front's scratch slot (several tiles: the last arriver adds them in tile order)
    double* slot = cx.bpart + ((int64_t)tk.w + rtile) * (KW * MR);
    {
        double* __restrict__ dst = ntiles > 1 ? slot : ps;
        const int ld = ntiles > 1 ? MR : MLD;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<double2*>(dst + (16 * w + 8 * i + fr) * ld + 8 * j + 2 * fc) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
    if (ntiles > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int old = atomicAdd(cx.counters2 + s, 1);
            s_last = ((old + 1) % ntiles) == 0;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        // tile-ordered sums: thread t owns elements t, t + 256, ... (16 of them); two tiles = 32 loads in flight per thread
        const double* __restrict__ base = cx.bpart + (int64_t)tk.w * (KW * MR) + tid;
        constexpr int NE = KW * MR / SOLVE_THREADS;
        double sum[NE];
#pragma unroll
        for (int u = 0; u < NE; ++u) sum[u] = 0.0;
        int t = 0;
        for (; t + 2 <= ntiles; t += 2) {
            double l0[NE], l1[NE];
#pragma unroll
            for (int u = 0; u < NE; ++u) {
                l0[u] = __ldcg(base + (int64_t)t * (KW * MR) + u * SOLVE_THREADS);
                l1[u] = __ldcg(base + (int64_t)(t + 1) * (KW * MR) + u * SOLVE_THREADS);
            }
#pragma unroll
            for (int u = 0; u < NE; ++u) sum[u] = (sum[u] + l0[u]) + l1[u];
        }
        if (t < ntiles) {
#pragma unroll
            for (int u = 0; u < NE; ++u) sum[u] += __ldcg(base + (int64_t)t * (KW * MR) + u * SOLVE_THREADS);
        }
#pragma unroll
        for (int u = 0; u < NE; ++u) {
            const int e = tid + u * SOLVE_THREADS;
            ps[(e >> 5) * MLD + (e & 31)] = sum[u];
        }
    }
    __syncthreads();
    // right-hand side of U11 X = x[cols] - part, in xs (reused); then the pivot block
    double* ys = xs;
#pragma unroll 8
    for (int a = w; a < KW; a += 8) ys[a * MLD + lane] = a < k ? x[(int64_t)(F.c0 + a) * MR + lane] - ps[a * MLD + lane] : 0.0;
    __syncthreads();
    pivot_solve32<false>(F, inv, ys);
    for (int a = w; a < k; a += 8) x[(int64_t)(F.c0 + a) * MR + lane] = ys[a * MLD + lane];
}

// ------------------------------------------------------------------ persistent chain solves
// A wide separator is stored as a chain of fronts (links) of at most KW pivot columns, each the only child of the next and
// its row list exactly the next link's front.  Level by level that is one launch per link and sweep -- hundreds of dependent
// launches of ~10-20 us for a 3D problem, an order of magnitude off the HBM time.  Here one persistent kernel runs a whole
// set of parallel chains (all CTAs co-resident: cooperative launch, one per SM): the chain's row space is cut into blocks
// (the pivot rows of link j = block j; the rows above the chain in blocks of 128) that are OWNED by CTAs round-robin.
// Forward, link j:  the owner of block j solves y_j = L11^-1 v_j and publishes it (solution vector + release flag);
// every CTA acquires it and applies v_B -= L21_j[B, :] y_j to the blocks B > j it owns, block j + 1 first -- its owner
// then solves and publishes y_{j+1} before touching its other blocks (look-ahead), so the dependent chain per link is
// flag -> one 128 x 128 product -> pivot solve, while the streaming of L21 trails behind on the other CTAs.
// Waits are bounded (~2 s): a lost producer raises the error flag instead of hanging the device.
__device__ __forceinline__ void chain_publish(int* flag, int epoch) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
__device__ __forceinline__ void chain_wait(const DevCtx& cx, const int* flag, int epoch) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int it = 0;; ++it) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v >= epoch) break;
            if ((it & 63) == 63) {
                // a producer that never shows up: raise the error flag once; every later wait sees it and gives up at once
                if (*(volatile int*)cx.flag == -4) break;
                if (clock64() - t0 > 4000000000LL) { atomicMin(cx.flag, -4); break; }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ const int* chain_of_cta(const int* __restrict__ chains, int nchains) {
    const int* d = chains;
    for (int q = 0; q < nchains; ++q, d += CHAIN_DESC)
        if ((int)blockIdx.x >= d[4] && (int)blockIdx.x < d[4] + d[5]) return d;
    return nullptr;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_fwd_chain(DevCtx cx, const int* __restrict__ chains, int nchains,
                                                                const double* __restrict__ win, double* __restrict__ zout, int epoch) {
    __shared__ double vown[CHAIN_MAXOWN][KW];
    __shared__ double ysA[KW], ysB[KW], red[KW];
    const int* d = chain_of_cta(chains, nchains);
    if (!d) return;
    const int* __restrict__ links = cx.chain_links + d[0];
    const int m = d[1], nblocks = d[3], c = (int)blockIdx.x - d[4], G = d[5];
    const int* __restrict__ boff = cx.chain_boff + d[2];
    const int tid = threadIdx.x, t = tid & (KW - 1), hf = tid >> 7;
    int* flags = cx.chain_flags;
    // geometry of every link, once (the per-link loop then has no dependent metadata loads on its critical path)
    __shared__ int lk_s[KW], lk_k[KW], lk_c0[KW];
    __shared__ int64_t lk_f[KW];
    __shared__ const double* lk_P[KW];
    __shared__ const double* lk_inv[KW];
    for (int j = tid; j < m; j += SOLVE_THREADS) {
        const int s = links[j];
        lk_s[j] = s; lk_c0[j] = cx.sn_start[s]; lk_k[j] = cx.sn_start[s + 1] - cx.sn_start[s];
        lk_f[j] = lk_k[j] + (cx.rows_ptr[s + 1] - cx.rows_ptr[s]);
        lk_P[j] = cx.lu + cx.Loff[s];
        lk_inv[j] = cx.dblk + (int64_t)cx.Doff[s] * (NB * NB);
    }
    int last_own = -1;
    for (int i = 0; c + i * G < nblocks; ++i) last_own = c + i * G;
    __syncthreads();
    // ---- right-hand side on the pivot blocks, zero above the chain
    for (int i = 0; c + i * G < nblocks; ++i) {
        const int b = c + i * G, nb = boff[b + 1] - boff[b];
        if (tid < KW) vown[i][tid] = (b < m && tid < nb) ? win[cx.sn_start[links[b]] + tid] : 0.0;
    }
    __syncthreads();
    // ---- the bottom link's children (any number, any index map), child by child in ascending order
    {
        const int s0 = links[0];
        for (int ci = cx.child_ptr[s0]; ci < cx.child_ptr[s0 + 1]; ++ci) {
            const int ch = cx.child_idx[ci];
            const int64_t rc = cx.rows_ptr[ch + 1] - cx.rows_ptr[ch];
            const int* __restrict__ rel = cx.rel + cx.rows_ptr[ch];
            const double* __restrict__ uc = cx.upd + cx.rows_ptr[ch];
            for (int64_t a = tid; a < rc; a += SOLVE_THREADS) {
                const int pos = rel[a];
                for (int i = 0; c + i * G < nblocks; ++i) {
                    const int b = c + i * G;
                    if (pos >= boff[b] && pos < boff[b + 1]) { vown[i][pos - boff[b]] += uc[a]; break; }
                }
            }
            __syncthreads();
        }
    }
    // pivot solve of link j from the block this CTA owns; the solution is left in ys and published
    auto solve = [&](int j, double* ys) {
        Front F;
        F.c0 = lk_c0[j]; F.k = lk_k[j]; F.f = lk_f[j]; F.r = F.f - F.k; F.P = const_cast<double*>(lk_P[j]); F.T = nullptr; F.C = nullptr;
        const double* __restrict__ inv = lk_inv[j];
        double y[1];
        y[0] = tid < F.k ? vown[j / G][tid] : 0.0;
        __syncthreads();
        diag_solve_lower<1, true>(F, inv, y, ys, tid);
        __syncthreads();
        if (tid < F.k) zout[F.c0 + tid] = ys[tid];
        chain_publish(flags + lk_s[j], epoch);
    };
    bool ahead = false;                       // y_j already solved (and sitting in ysB) by the look-ahead of step j - 1
    for (int j = 0; j < m && j <= last_own; ++j) {         // (a CTA whose blocks are all behind the chain's front is done)
        const int s = lk_s[j];
        struct { int k, c0; int64_t f; const double* P; } F = {lk_k[j], lk_c0[j], lk_f[j], lk_P[j]};
        if (j % G == c) {
            if (ahead) { if (tid < KW) ysA[tid] = ysB[tid]; __syncthreads(); }
            else solve(j, ysA);
        } else {
            chain_wait(cx, flags + s, epoch);
            if (tid < KW) ysA[tid] = tid < F.k ? __ldcg(zout + F.c0 + tid) : 0.0;
            __syncthreads();
        }
        ahead = false;
        // blocks B > j this CTA owns, block j + 1 first
        const int first = j + 1;
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; c + i * G < nblocks; ++i) {
                const int b = c + i * G;
                if (b <= j || (pass == 0) != (b == first)) continue;
                const int nb = boff[b + 1] - boff[b];
                // rows of block b inside P_j: below the pivot block, offset by the chain rows between
                const double* __restrict__ src = F.P + F.k + (boff[b] - boff[j + 1]) + t + (int64_t)(hf * (KW / 2)) * F.f;
                const int jn = F.k - hf * (KW / 2) < KW / 2 ? F.k - hf * (KW / 2) : KW / 2;      // columns of this half
                double a4[4] = {0.0, 0.0, 0.0, 0.0};
                if (t < nb) {
                    const double* __restrict__ yh = ysA + hf * (KW / 2);
                    int u0 = 0;
                    for (; u0 + 32 <= jn; u0 += 32) {
                        double l[32];
#pragma unroll
                        for (int u = 0; u < 32; ++u) l[u] = src[(int64_t)(u0 + u) * F.f];
#pragma unroll
                        for (int u = 0; u < 32; ++u) a4[u & 3] += l[u] * yh[u0 + u];
                    }
                    for (; u0 < jn; ++u0) a4[u0 & 3] += src[(int64_t)u0 * F.f] * yh[u0];
                }
                const double part = (a4[0] + a4[1]) + (a4[2] + a4[3]);
                if (hf == 1) red[t] = part;
                __syncthreads();
                if (hf == 0 && t < nb) vown[i][t] -= part + red[t];
                __syncthreads();
                if (pass == 0 && first < m) { solve(first, ysB); ahead = true; }
            }
    }
    // ---- what is left on the rows above the chain is the top link's update vector
    {
        const int sl = links[m - 1];
        double* __restrict__ us = cx.upd + cx.rows_ptr[sl];
        for (int i = 0; c + i * G < nblocks; ++i) {
            const int b = c + i * G;
            if (b < m) continue;
            const int nb = boff[b + 1] - boff[b];
            if (tid < nb) us[boff[b] - boff[m] + tid] = vown[i][tid];
        }
    }
}

// Backward, rows above the chain: partial products  part[link][tile][c] = sum_{a in tile} U12'_link[a, c] x[rows[a]]  for every
// link of every chain of the set at once (x above the chain is final: no dependencies).  task: x = link, y = first row of the
// tile inside T_link, z = rows, w = slot (KW-vector) of the result.
__global__ void __launch_bounds__(SOLVE_THREADS, 2) k_bwd_rect(DevCtx cx, const int4* __restrict__ tasks, const double* __restrict__ x) {
    __shared__ double xs[BWD_ROWS];
    const int4 tk = tasks[blockIdx.x];
    const Front F = load_front(cx, tk.x);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cnt = tk.z;
    const int* __restrict__ rows = cx.rows + cx.rows_ptr[tk.x] + tk.y;
    xs[tid] = tid < cnt ? x[rows[tid]] : 0.0;
    __syncthreads();
    double* __restrict__ out = cx.chain_part + (int64_t)tk.w * KW;
    constexpr int BU = 4;
    for (int i0 = warp * BU; i0 < F.k; i0 += (SOLVE_THREADS / 32) * BU) {
        double tv[BU][BWD_ROWS / 32], v[BU];
#pragma unroll
        for (int u = 0; u < BU; ++u) {
            const double* __restrict__ col = F.T + (int64_t)(i0 + u) * F.r + tk.y;
#pragma unroll
            for (int a = 0; a < BWD_ROWS / 32; ++a) tv[u][a] = (i0 + u < F.k && lane + 32 * a < cnt) ? col[lane + 32 * a] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < BU; ++u) {
            v[u] = 0.0;
#pragma unroll
            for (int a = 0; a < BWD_ROWS / 32; ++a) v[u] += tv[u][a] * xs[lane + 32 * a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
            if (lane == 0 && i0 + u < F.k) out[i0 + u] = v[u];
        }
    }
}

// Backward chain: link i = m-1 .. 0.  The owner of link i solves x_i = U11^-1 (pending_i) and publishes it; every CTA
// applies  pending_j -= U[j-block, i-block] x_i  to the links j < i it owns (the block is rows [o_i - o_{j+1}, +k_i) of T_j),
// link i - 1 first, whose owner then solves and publishes x_{i-1} before its other links (look-ahead).
__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_bwd_chain(DevCtx cx, const int* __restrict__ chains, int nchains,
                                                                double* __restrict__ x, int epoch) {
    __shared__ double pend[CHAIN_MAXOWN][KW];
    __shared__ double xsA[KW], xsB[KW];
    const int* d = chain_of_cta(chains, nchains);
    if (!d) return;
    const int* __restrict__ links = cx.chain_links + d[0];
    const int m = d[1], c = (int)blockIdx.x - d[4], G = d[5], ntile = d[7];
    const int* __restrict__ boff = cx.chain_boff + d[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* flags = cx.chain_flags + cx.chain_nsn;
    // link j is owned by CTA (m - 1 - j) % G, slot (m - 1 - j) / G
    auto owner = [&](int j) { return (m - 1 - j) % G; };
    auto slot = [&](int j) { return (m - 1 - j) / G; };
    // ---- pending right-hand sides: forward solution minus the (tile-ordered) products with the rows above the chain
    for (int j = m - 1 - c; j >= 0; j -= G) {
        const int s = links[j], k = cx.sn_start[s + 1] - cx.sn_start[s];
        if (tid < KW) {
            double v = tid < k ? x[cx.sn_start[s] + tid] : 0.0;
            const double* __restrict__ pp = cx.chain_part + ((int64_t)d[6] + (int64_t)j * ntile) * KW + tid;
            for (int tl = 0; tl < ntile; ++tl) v -= (tid < k ? pp[(int64_t)tl * KW] : 0.0);
            pend[slot(j)][tid] = v;
        }
    }
    __syncthreads();
    auto solve = [&](int i, double* xs) {
        const int s = links[i];
        const Front F = load_front(cx, s);
        const double* __restrict__ inv = cx.dblk + (int64_t)cx.Doff[s] * (NB * NB);
        double v[1];
        v[0] = tid < F.k ? pend[slot(i)][tid] : 0.0;
        __syncthreads();
        diag_solve_upper<1, true>(F, inv, v, xs, tid);
        __syncthreads();
        if (tid < F.k) x[F.c0 + tid] = xs[tid];
        chain_publish(flags + s, epoch);
    };
    bool ahead = false;
    for (int i = m - 1; i >= 0; --i) {
        const int s = links[i], ki = cx.sn_start[s + 1] - cx.sn_start[s];
        if (owner(i) == c) {
            if (ahead) { if (tid < KW) xsA[tid] = xsB[tid]; __syncthreads(); }
            else solve(i, xsA);
        } else {
            chain_wait(cx, flags + s, epoch);
            if (tid < KW) xsA[tid] = tid < ki ? __ldcg(x + cx.sn_start[s] + tid) : 0.0;
            __syncthreads();
        }
        ahead = false;
        const int first = i - 1;
        for (int pass = 0; pass < 2; ++pass)
            for (int j = m - 1 - c; j >= 0; j -= G) {
                if (j >= i || (pass == 0) != (j == first)) continue;
                const Front F = load_front(cx, links[j]);
                // rows of link i's pivots inside T_j (r_j x k_j): a warp takes columns c0 = warp, warp + 8, ...; lanes = rows
                const double* __restrict__ blk = F.T + (boff[i] - boff[j + 1]);
                double xv[KW / 32];
#pragma unroll
                for (int a = 0; a < KW / 32; ++a) xv[a] = xsA[lane + 32 * a];        // zero beyond k_i
                for (int c0 = warp; c0 < F.k; c0 += 4 * (SOLVE_THREADS / 32)) {
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int cc = c0 + u * (SOLVE_THREADS / 32);
                        const double* __restrict__ col = blk + (int64_t)cc * F.r;
                        v[u] = 0.0;
                        if (cc < F.k) {
#pragma unroll
                            for (int a = 0; a < KW / 32; ++a) v[u] += (lane + 32 * a < ki ? col[lane + 32 * a] : 0.0) * xv[a];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
                        const int cc = c0 + u * (SOLVE_THREADS / 32);
                        if (lane == 0 && cc < F.k) pend[slot(j)][cc] -= v[u];
                    }
                }
                __syncthreads();
                if (pass == 0 && first >= 0) { solve(first, xsB); ahead = true; }
            }
    }
}

// ------------------------------------------------------------------ partitioned solves
// One CTA per interface front: sum the update vectors of this rank's subtree roots below it into the
// front's virtual child (rel = identity), root by root in ascending order.
// task: x = interface front, y = virtual child, [z, w) = range in vlist.
template <int RB>
__global__ void __launch_bounds__(256) k_vgather(DevCtx cx, const int4* __restrict__ tasks, const int* __restrict__ vlist) {
    const int4 tk = tasks[blockIdx.x];
    double* __restrict__ v = cx.upd + cx.rows_ptr[tk.y] * RB;
    const int64_t fs = cx.rows_ptr[tk.y + 1] - cx.rows_ptr[tk.y];
    for (int64_t i = threadIdx.x; i < fs * RB; i += 256) v[i] = 0.0;
    __syncthreads();
    for (int ci = tk.z; ci < tk.w; ++ci) {
        const int c = vlist[ci];
        const int64_t rc = cx.rows_ptr[c + 1] - cx.rows_ptr[c];
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ uc = cx.upd + cx.rows_ptr[c] * RB;
        for (int64_t a = threadIdx.x; a < rc; a += 256) {
            const int64_t ra = (int64_t)rel[a] * RB;
#pragma unroll
            for (int q = 0; q < RB; ++q) v[ra + q] += uc[a * RB + q];
        }
        __syncthreads();
    }
}
// keep the entries this rank is responsible for (its own columns; the top columns on rank 0)
template <int RB>
__global__ void k_mask_owned(int n, const int* __restrict__ colowner, int rank, double* __restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = colowner[i];
    if (!(o == rank || (o == -1 && rank == 0))) {
#pragma unroll
        for (int q = 0; q < RB; ++q) z[(int64_t)i * RB + q] = 0.0;
    }
}

// ------------------------------------------------------------------ partitioned top: replication and sync
// Copy segments [off, off + len) (doubles, both even) of this rank's factor pool to the same offsets of every
// peer's pool: the panel owner publishes P_s (pivot block + L21) of a top front.  One read, nranks - 1 remote
// writes over NVLink, 16 bytes per lane.  segs: (off, len) pairs; blockIdx.y = segment.
__global__ void __launch_bounds__(256) k_replicate(DevCtx cx, const int64_t* __restrict__ segs) {
    const int64_t off = segs[2 * blockIdx.y], len2 = segs[2 * blockIdx.y + 1] >> 1;
    const double2* __restrict__ src = reinterpret_cast<const double2*>(cx.lu + off);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < len2; i += (int64_t)gridDim.x * 256) {
        const double2 v = src[i];
        for (int g = 0; g < cx.nranks; ++g)
            if (g != cx.rank) reinterpret_cast<double2*>(cx.lu_peer[g] + off)[i] = v;
    }
}
// Rows of U12' (T_s, r x k, column-major) this rank owns and computed, to every peer: the solves read the whole of T_s
// on every rank.  task: x = supernode, y = first row, z = number of rows (a run with one owner), one CTA per task.
__global__ void __launch_bounds__(256) k_replicate_rows(DevCtx cx, const int4* __restrict__ tasks) {
    const int4 tk = tasks[blockIdx.x];
    const int s = tk.x, k = cx.sn_start[s + 1] - cx.sn_start[s];
    const int64_t r = cx.rows_ptr[s + 1] - cx.rows_ptr[s], off = cx.Uoff[s] + tk.y;
    for (int e = threadIdx.x; e < tk.z * k; e += 256) {
        const int c = e / tk.z, a = e - c * tk.z;
        const int64_t o = off + a + (int64_t)c * r;
        const double v = cx.lu[o];
        for (int g = 0; g < cx.nranks; ++g)
            if (g != cx.rank) cx.lu_peer[g][o] = v;
    }
}
// Cross-GPU signalling (each rank drives its own GPU, so a waiting kernel never shares a device with its producer):
// k_signal publishes `epoch` in slot `slot` of every peer's flag array after making this rank's earlier writes visible
// system-wide; k_wait spins on this rank's own flags.  A wait that lasts longer than ~4 s gives up and raises the
// pivot flag to -3, so that a lost peer shows up as an error instead of a hang.
__global__ void k_signal(DevCtx cx, int slot, int epoch) {
    __threadfence_system();
    const int g = threadIdx.x;
    if (g < cx.nranks && g != cx.rank) {
        int* f = cx.xflag_peer[g] + slot;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
}
__global__ void k_wait(DevCtx cx, const int* __restrict__ slots, int nslots, int epoch) {
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) {
        const int* f = cx.xflag_peer[cx.rank] + slots[i];
        const long long t0 = clock64();
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 8000000000LL) { atomicMin(cx.flag, -3); break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

// ------------------------------------------------------------------ solves, small fronts
// One warp per front (k <= 32, f <= SMALL_F_MAX), FPC fronts per CTA, no block-wide barriers.
// Forward: v = [w[cols]; 0] + children's update vectors (gathered through rel, child by child),
// then column-oriented substitution: for j < k: y_j = v_j (broadcast by shuffle), v_i -= L_ij y_j
// for all i > j -- rows of L11 and of L21 alike, each column of P read once, coalesced.
// LDV = slots of the interleaved work vectors (RB, or 32 when the sweep is one of the four 8-wide slices of a
// 32-wide sweep: blockIdx.y picks the slice)
template <int FPC, int RB, int LDV = RB>
__global__ void __launch_bounds__(32 * FPC) k_small_fwd(DevCtx cx, const int4* __restrict__ tasks, int ntasks,
                                                        const double* __restrict__ win, double* __restrict__ zout) {
    __shared__ double vs[FPC][SMALL_F_MAX * RB];
    const int grp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int ti = blockIdx.x * FPC + grp;
    if (ti >= ntasks) return;
    pdl_trigger();
    const int s = tasks[ti].x;
    const Front F = load_front(cx, s);
    pdl_wait();
    const int k = F.k, f = (int)F.f, q0 = LDV == RB ? 0 : (int)blockIdx.y * RB;
    double* v = vs[grp];
    for (int e = lane; e < f * RB; e += 32) v[e] = e < k * RB ? win[(int64_t)(F.c0 + e / RB) * LDV + q0 + (e % RB)] : 0.0;
    __syncwarp();
    for (int ci = cx.child_ptr[s]; ci < cx.child_ptr[s + 1]; ++ci) {
        const int c = cx.child_idx[ci];
        const int rc = (int)(cx.rows_ptr[c + 1] - cx.rows_ptr[c]);
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ uc = cx.upd + cx.rows_ptr[c] * LDV + q0;
        for (int a = lane; a < rc; a += 32) {
            const int ra = rel[a] * RB;
#pragma unroll
            for (int q = 0; q < RB; ++q) v[ra + q] += uc[(int64_t)a * LDV + q];
        }
        __syncwarp();
    }
    const bool h0 = lane < f, h1 = lane + 32 < f, h2 = lane + 64 < f;
    double v0[RB], v1[RB], v2[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
        v0[q] = h0 ? v[lane * RB + q] : 0.0;
        v1[q] = h1 ? v[(lane + 32) * RB + q] : 0.0;
        v2[q] = h2 ? v[(lane + 64) * RB + q] : 0.0;
    }
    const double* __restrict__ col = F.P + lane;
#pragma unroll 4
    for (int j = 0; j < k; ++j, col += f) {
        const double l0 = (h0 && lane > j) ? col[0] : 0.0, l1 = h1 ? col[32] : 0.0, l2 = h2 ? col[64] : 0.0;
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const double yj = __shfl_sync(0xffffffffu, v0[q], j);
            v0[q] -= l0 * yj; v1[q] -= l1 * yj; v2[q] -= l2 * yj;
        }
    }
    double* __restrict__ us = cx.upd + (cx.rows_ptr[s] - k) * LDV + q0;
#pragma unroll
    for (int q = 0; q < RB; ++q) {
        if (lane < k) zout[(int64_t)(F.c0 + lane) * LDV + q0 + q] = v0[q];
        if (h0 && lane >= k) us[lane * LDV + q] = v0[q];
        if (h1) us[(lane + 32) * LDV + q] = v1[q];                  // k <= 32 <= lane + 32
        if (h2) us[(lane + 64) * LDV + q] = v2[q];
    }
}

// Backward: lane c owns pivot row c.  t_c = sum_b U12[c][b] x[rows[b]] is a plain loop over b (the
// r x k block is a few KB and stays in L1, so the strided reads cost nothing in HBM traffic and no
// reduction is needed); then column-oriented back substitution with U11 and the stored 1/u_jj.
template <int FPC, int RB, int LDV = RB>
__global__ void __launch_bounds__(32 * FPC) k_small_bwd(DevCtx cx, const int4* __restrict__ tasks, int ntasks,
                                                        double* __restrict__ x) {
    __shared__ double xs[FPC][SMALL_F_MAX * RB];
    const int grp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int ti = blockIdx.x * FPC + grp;
    if (ti >= ntasks) return;
    pdl_trigger();
    const int s = tasks[ti].x;
    const Front F = load_front(cx, s);
    pdl_wait();
    const int k = F.k, r = (int)F.r, f = (int)F.f, q0 = LDV == RB ? 0 : (int)blockIdx.y * RB;
    double* xr = xs[grp];
    const int* __restrict__ rows = cx.rows + cx.rows_ptr[s];
    for (int b = lane; b < r; b += 32) {
        const int64_t xo = (int64_t)rows[b] * LDV + q0;
#pragma unroll
        for (int q = 0; q < RB; ++q) xr[b * RB + q] = x[xo + q];
    }
    __syncwarp();
    const bool act = lane < k;
    double v[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) v[q] = 0.0;
    {
        const double* __restrict__ tc = F.T + (int64_t)(act ? lane : 0) * r;
#pragma unroll 4
        for (int b = 0; b < r; ++b) {
            const double tv = tc[b];
#pragma unroll
            for (int q = 0; q < RB; ++q) v[q] += tv * xr[b * RB + q];
        }
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) v[q] = act ? x[(int64_t)(F.c0 + lane) * LDV + q0 + q] - v[q] : 0.0;
    const double d = act ? cx.dinv[F.c0 + lane] : 0.0;
    const double* __restrict__ col = F.P + lane + (int64_t)(k - 1) * f;
#pragma unroll 4
    for (int j = k - 1; j >= 0; --j, col -= f) {
        const double u = lane < j ? col[0] : 0.0;
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const double xj = __shfl_sync(0xffffffffu, v[q] * d, j);
            v[q] = lane == j ? xj : v[q] - u * xj;
        }
    }
    if (act) {
#pragma unroll
        for (int q = 0; q < RB; ++q) x[(int64_t)(F.c0 + lane) * LDV + q0 + q] = v[q];
    }
}

}  // namespace

int front_small_limit() { return SMALL_F_MAX; }

int debug_read_trace(long long* out) {
#ifdef SMSLU_TRACE
    return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 32) == cudaSuccess ? 32 : -1;
#else
    (void)out;
    return 0;
#endif
}

constexpr int SMALL_FPC32 = 4;     // fronts per CTA in the one-warp class of k_small_factor

static size_t gemm_smem() { return sizeof(double) * GEMM_STAGES * 2 * GKC * GEMM_LDS; }
static size_t gemm_strip_smem() { return sizeof(double) * (KW + 2 * NB) * GEMM_LDS; }
static size_t panel_smem(int j0, int rows) {
    return sizeof(double) * ((2 * (size_t)j0 + rows) * CLD + (rows == PANEL_ROWS_TOP ? (size_t)rows * (KW + 4) : 0));
}

// launch with programmatic stream serialization (see pdl_trigger / pdl_wait) when the grid is small
template <class... KArgs, class... Args>
static void launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = !(getenv("SMSLU_NO_PDL") && atoi(getenv("SMSLU_NO_PDL")) != 0);   // debugging aid
    // only where latency matters: early-launched CTAs of a big dependent grid sit on SM resources while the
    // predecessor's tail is still running (measured: +13 % on a 3D 96^3 refactorization)
    cfg.attrs = attr; cfg.numAttrs = (pdl && grid <= 4 * 148) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

cudaError_t kernels_init() {
    cudaError_t e = cudaFuncSetAttribute(k_small_factor<96, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(double) * small_group_doubles(96)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_gemm_cb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fwd32<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (KW + MROWS) * MLD));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fwd32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (KW + RB_WIDE_ROWS_TOP) * MLD));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bwd32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (KW + MROWS) * MLD));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_gemm_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_strip_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_panel<PANEL_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem(KW - NB, PANEL_ROWS));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_panel<PANEL_ROWS_TOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem(KW - NB, PANEL_ROWS_TOP));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_assemble_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * ASM_COLS * ASM_SMEM_ROWS));
    if (e != cudaSuccess) return e;
    return cudaSuccess;
}

void launch_diag_inverse(cudaStream_t st, const DevCtx& cx, const int2* tasks, int ntasks) {
    if (ntasks > 0) k_diag_inverse<<<(ntasks + INV_WARPS - 1) / INV_WARPS, 32 * INV_WARPS, 0, st>>>(cx, tasks, ntasks);
}
void launch_rowscale(cudaStream_t st, int n, const int64_t* rowptr, const int64_t* rowidx, const double* av, double* Rs) {
    k_rowscale<<<(n + 255) / 256, 256, 0, st>>>(n, rowptr, rowidx, av, Rs);
}
void launch_scatter(cudaStream_t st, int64_t nnz, const int64_t* dst, const int* arow, const int* asrc,
                    const double* Rs, const double* av, double* lu) {
    if (nnz <= 0) return;
    int64_t blocks = (nnz + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_scatter<<<(int)blocks, 256, 0, st>>>(nnz, dst, arow, asrc, Rs, av, lu);
}
void launch_zero_cb(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) launch_pdl(k_zero_cb, ntasks, 256, 0, st, cx, tasks);
}
void launch_assemble(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int fmax) {
    if (ntasks <= 0) return;
    if (fmax > 0) launch_pdl(k_assemble_smem, ntasks, 256, sizeof(double) * ASM_COLS * (fmax < ASM_SMEM_ROWS ? fmax : ASM_SMEM_ROWS), st, cx, tasks);
    else launch_pdl(k_assemble, ntasks, 256, 0, st, cx, tasks);       // (never: every assembly launch knows its largest parent)
}
template <int RW, int CH, int FPC>
static void launch_small_class(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int fmax,
                               const double* av, const double* Rs) {
    launch_pdl(k_small_factor<RW, CH, FPC>, (ntasks + FPC - 1) / FPC, RW * CH * FPC,
               sizeof(double) * small_group_doubles(fmax) * FPC, st, cx, tasks, ntasks, fmax, av, Rs);
}
void launch_front_small(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int fmax,
                        const double* av, const double* Rs) {
    if (ntasks <= 0) return;
    if (fmax <= 32)
        launch_pdl(k_small_factor_reg<32, SMALL_FPC32>, (ntasks + SMALL_FPC32 - 1) / SMALL_FPC32, 32 * SMALL_FPC32,
                   sizeof(double) * reg_group_doubles(32) * SMALL_FPC32, st, cx, tasks, ntasks, av, Rs);
    else if (fmax <= 40)
        launch_pdl(k_small_factor_reg<40, 1>, ntasks, 64, sizeof(double) * reg_group_doubles(40), st, cx, tasks, ntasks, av, Rs);
    else if (fmax <= 48)
        launch_pdl(k_small_factor_reg<48, 1>, ntasks, 64, sizeof(double) * reg_group_doubles(48), st, cx, tasks, ntasks, av, Rs);
    else if (fmax <= 64)
        launch_pdl(k_small_factor_reg<64, 1>, ntasks, 64, sizeof(double) * reg_group_doubles(64), st, cx, tasks, ntasks, av, Rs);
    else launch_small_class<96, 3, 1>(st, cx, tasks, ntasks, fmax, av, Rs);
}
void launch_replicate(cudaStream_t st, const DevCtx& cx, const int64_t* segs, int nsegs) {
    if (nsegs > 0) k_replicate<<<dim3(148, nsegs), 256, 0, st>>>(cx, segs);
}
void launch_replicate_rows(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) k_replicate_rows<<<ntasks, 256, 0, st>>>(cx, tasks);
}
void launch_signal(cudaStream_t st, const DevCtx& cx, int slot, int epoch) { k_signal<<<1, 32, 0, st>>>(cx, slot, epoch); }
void launch_wait(cudaStream_t st, const DevCtx& cx, const int* slots, int nslots, int epoch) {
    if (nslots > 0) k_wait<<<1, 32, 0, st>>>(cx, slots, nslots, epoch);
}
#define RB_DISPATCH(rb, CALL) do { if ((rb) == 1) { CALL(1); } else if ((rb) == 4) { CALL(4); } else { CALL(8); } } while (0)
void launch_vgather(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const int* vlist, int rb) {
    if (ntasks <= 0) return;
#define CALL(R) k_vgather<R><<<ntasks, 256, 0, st>>>(cx, tasks, vlist)
    RB_DISPATCH(rb, CALL);
#undef CALL
}
void launch_mask_owned(cudaStream_t st, int n, const int* colowner, int rank, double* z, int rb) {
#define CALL(R) k_mask_owned<R><<<(n + 255) / 256, 256, 0, st>>>(n, colowner, rank, z)
    RB_DISPATCH(rb, CALL);
#undef CALL
}
void launch_small_fwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const double* win, double* zout, int rb) {
    if (ntasks <= 0) return;
    if (rb == MR) { k_small_fwd<2, 8, MR><<<dim3((ntasks + 1) / 2, MR / 8), 64, 0, st>>>(cx, tasks, ntasks, win, zout); return; }
    if (rb == 1) launch_pdl(k_small_fwd<8, 1>, (ntasks + 7) / 8, 256, 0, st, cx, tasks, ntasks, win, zout);
    else if (rb == 4) launch_pdl(k_small_fwd<4, 4>, (ntasks + 3) / 4, 128, 0, st, cx, tasks, ntasks, win, zout);
    else launch_pdl(k_small_fwd<2, 8>, (ntasks + 1) / 2, 64, 0, st, cx, tasks, ntasks, win, zout);
}
void launch_small_bwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x, int rb) {
    if (ntasks <= 0) return;
    if (rb == MR) { k_small_bwd<2, 8, MR><<<dim3((ntasks + 1) / 2, MR / 8), 64, 0, st>>>(cx, tasks, ntasks, x); return; }
    if (rb == 1) launch_pdl(k_small_bwd<8, 1>, (ntasks + 7) / 8, 256, 0, st, cx, tasks, ntasks, x);
    else if (rb == 4) launch_pdl(k_small_bwd<4, 4>, (ntasks + 3) / 4, 128, 0, st, cx, tasks, ntasks, x);
    else launch_pdl(k_small_bwd<2, 8>, (ntasks + 1) / 2, 64, 0, st, cx, tasks, ntasks, x);
}
void launch_panel(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int g, int rows) {
    if (ntasks <= 0) return;
    if (rows == PANEL_ROWS) launch_pdl(k_panel<PANEL_ROWS>, ntasks, PANEL_THREADS, panel_smem(g * NB, PANEL_ROWS), st, cx, tasks);
    else launch_pdl(k_panel<PANEL_ROWS_TOP>, ntasks, PANEL_THREADS, panel_smem(g * NB, PANEL_ROWS_TOP), st, cx, tasks);
}
void launch_gemm_cb(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) launch_pdl(k_gemm_cb, ntasks, 256, gemm_smem(), st, cx, tasks, ntasks);
}
void launch_gemm_strip(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) launch_pdl(k_gemm_strip, ntasks, 256, gemm_strip_smem(), st, cx, tasks);
}
template <class... KArgs>
static cudaError_t launch_coop(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, KArgs... args) {
    void* argv[] = {(void*)&args...};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(block), argv, 0, st);
}
cudaError_t launch_fwd_chain(cudaStream_t st, const DevCtx& cx, const int* chains, int nchains, int nctas, const double* win, double* zout, int epoch) {
    return launch_coop(k_fwd_chain, nctas, SOLVE_THREADS, st, cx, chains, nchains, win, zout, epoch);
}
void launch_bwd_rect(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const double* x) {
    if (ntasks > 0) k_bwd_rect<<<ntasks, SOLVE_THREADS, 0, st>>>(cx, tasks, x);
}
cudaError_t launch_bwd_chain(cudaStream_t st, const DevCtx& cx, const int* chains, int nchains, int nctas, double* x, int epoch) {
    return launch_coop(k_bwd_chain, nctas, SOLVE_THREADS, st, cx, chains, nchains, x, epoch);
}
void launch_permute_scale(cudaStream_t st, int n, const int* p, const double* Rs, const double* b, int64_t ldb, double* w, int rb, int nv) {
#define CALL(R) k_permute_scale<R><<<(n + 255) / 256, 256, 0, st>>>(n, p, Rs, b, ldb, w, nv)
    if (rb == MR) { CALL(MR); return; }
    RB_DISPATCH(rb, CALL);
#undef CALL
}
void launch_unpermute(cudaStream_t st, int n, const int* q, const double* w, double* x, int64_t ldx, int rb, int nv) {
#define CALL(R) k_unpermute<R><<<(n + 255) / 256, 256, 0, st>>>(n, q, w, x, ldx, nv)
    if (rb == MR) { CALL(MR); return; }
    RB_DISPATCH(rb, CALL);
#undef CALL
}
void launch_fwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int rows, const double* win, double* zout, int rb) {
    if (ntasks <= 0) return;
    if (rows == FWD_ROWS) {
#define CALL(R) launch_pdl(k_fwd<R, FWD_ROWS>, ntasks, SOLVE_THREADS, 0, st, cx, tasks, win, zout)
        RB_DISPATCH(rb, CALL);
#undef CALL
    } else if (rows == FWD_ROWS_MID) {
#define CALL(R) launch_pdl(k_fwd<R, FWD_ROWS_MID>, ntasks, SOLVE_THREADS, 0, st, cx, tasks, win, zout)
        RB_DISPATCH(rb, CALL);
#undef CALL
    } else {
#define CALL(R) launch_pdl(k_fwd<R, FWD_ROWS_WIDE>, ntasks, SOLVE_THREADS, 0, st, cx, tasks, win, zout)
        RB_DISPATCH(rb, CALL);
#undef CALL
    }
}
static size_t mrhs_smem() { return sizeof(double) * (KW + MROWS) * MLD; }
void launch_fwd32(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int rows, const double* win, double* zout) {
    if (ntasks <= 0) return;
    if (rows == RB_WIDE_ROWS) launch_pdl(k_fwd32<32>, ntasks, SOLVE_THREADS, mrhs_smem(), st, cx, tasks, win, zout);
    else launch_pdl(k_fwd32<8>, ntasks, SOLVE_THREADS, sizeof(double) * (KW + RB_WIDE_ROWS_TOP) * MLD, st, cx, tasks, win, zout);
}
void launch_bwd32(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x) {
    if (ntasks > 0) launch_pdl(k_bwd32, ntasks, SOLVE_THREADS, mrhs_smem(), st, cx, tasks, x);
}
void launch_bwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x, int rb) {
    if (ntasks <= 0) return;
    if (ntasks <= BWD_WIDE_TILES) {
#define CALL(R) launch_pdl(k_bwd<R, false>, ntasks, SOLVE_THREADS, 0, st, cx, tasks, x)
        RB_DISPATCH(rb, CALL);
#undef CALL
    } else {                                                 // throughput variant: two CTAs per SM
#define CALL(R) launch_pdl(k_bwd<R, true>, ntasks, SOLVE_THREADS, 0, st, cx, tasks, x)
        RB_DISPATCH(rb, CALL);
#undef CALL
    }
}

}  // namespace smslu
