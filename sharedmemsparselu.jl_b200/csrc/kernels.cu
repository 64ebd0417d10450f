// sm_100a kernels, see kernels.cuh for the data layout.  FP64 throughout.
#include "kernels.cuh"

#include <math.h>

namespace smslu {

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

__device__ __forceinline__ bool bad_pivot(double p) { return !(fabs(p) > 0.0) || !isfinite(p); }

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
__global__ void k_scatter(int64_t nnz, const int64_t* __restrict__ dst, const int* __restrict__ arow,
                          const double* __restrict__ Rs, const double* __restrict__ av, double* __restrict__ lu) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride)
        lu[dst[t]] = Rs[arow[t]] * av[t];
}

// ------------------------------------------------------------------ zero contribution blocks
// task: x = supernode, y = tile
__global__ void __launch_bounds__(256) k_zero_cb(DevCtx cx, const int4* __restrict__ tasks) {
    int4 tk = tasks[blockIdx.x];
    int64_t r = cx.rows_ptr[tk.x + 1] - cx.rows_ptr[tk.x];
    int64_t tot = r * r;
    double* C = cx.cb + cx.CBoff[tk.x];
    int64_t lo = (int64_t)tk.y * ZERO_TILE, hi = lo + ZERO_TILE;
    if (hi > tot) hi = tot;
    for (int64_t e = lo + threadIdx.x; e < hi; e += 256) C[e] = 0.0;
}

// ------------------------------------------------------------------ extend-add child CB into parent
// task: x = child supernode, y = first child-CB column, z = number of columns.
// Within one launch every parent receives from at most one child => no write conflicts and a
// fixed summation order (children are applied slot by slot, in ascending child order).
__global__ void __launch_bounds__(256) k_extend_add(DevCtx cx, const int4* __restrict__ tasks) {
    int4 tk = tasks[blockIdx.x];
    const int c = tk.x;
    const int64_t rc = cx.rows_ptr[c + 1] - cx.rows_ptr[c];
    const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
    const double* __restrict__ Cc = cx.cb + cx.CBoff[c];
    const Front F = load_front(cx, cx.sn_parent[c]);
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int b = tk.y + ty; b < tk.y + tk.z; b += 4) {
        const int rb = rel[b];
        const double* __restrict__ src = Cc + (int64_t)b * rc;
        if (rb < F.k) {
            double* dst = F.P + (int64_t)rb * F.f;
            for (int a = tx; a < rc; a += 64) dst[rel[a]] += src[a];
        } else {
            double* dstC = F.C + (int64_t)(rb - F.k) * F.r - F.k;
            double* dstT = F.T + (rb - F.k);
            for (int a = tx; a < rc; a += 64) {
                const int ra = rel[a];
                if (ra < F.k) dstT[(int64_t)ra * F.r] += src[a];
                else dstC[ra] += src[a];
            }
        }
    }
}

// ------------------------------------------------------------------ warp-level 32x32 LU
// One warp factors a k x k (k <= 32) block held one row per lane in registers, in the given
// (static) pivot order: on exit lane i holds row i of the packed factors (L strictly below the
// diagonal, U on and above).  No shared memory, no block barriers; the pivot row travels by
// shuffles.  Entries with row or column >= k must be zero on entry.
__device__ __forceinline__ void warp_lu32(double (&a)[KMAX], int lane, int k, int c0, int* flag) {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j >= k) break;
        const double piv = __shfl_sync(0xffffffffu, a[j], j);
        if (lane == 0 && bad_pivot(piv)) atomicMin(flag, c0 + j);
        double l = 0.0;
        if (lane > j) { l = a[j] / piv; a[j] = l; }
#pragma unroll
        for (int c = j + 1; c < KMAX; ++c) {
            const double u = __shfl_sync(0xffffffffu, a[c], j);
            if (lane > j) a[c] -= l * u;
        }
    }
}

// Row of L21 against U11 (x <- x U11^{-1}) or row of U12' against L11' (x <- x L11^{-T}, unit
// diagonal); D holds the packed factors of the pivot block in shared memory.
__device__ __forceinline__ void trsm_row_upper(double (&x)[KMAX], const double (*D)[KMAX + 1], int k) {
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
        if (c >= k) break;
        double v = x[c];
#pragma unroll
        for (int p = 0; p < c; ++p) v -= x[p] * D[p][c];
        x[c] = v / D[c][c];
    }
}
__device__ __forceinline__ void trsm_row_lower_t(double (&x)[KMAX], const double (*D)[KMAX + 1], int k) {
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
        if (c >= k) break;
        double v = x[c];
#pragma unroll
        for (int p = 0; p < c; ++p) v -= x[p] * D[c][p];
        x[c] = v;
    }
}

// ------------------------------------------------------------------ fused small front
// One CTA does a whole front with f <= SMALL_F_MAX: warp 0 factors the pivot block in registers
// (warp_lu32) while the other warps stage L21 and U12' in shared memory; then every thread solves
// rows of the two panels against the pivot block, and the CTA forms the contribution block
// C = beta*C - L21 U12 from shared memory (4x4 micro-tiles) and streams it to HBM.
// task: x = supernode, y = beta (the block holds assembled child contributions).
// dynamic shared memory: 2 * KMAX * rp doubles, rp = r rounded up to 4 (+4 padding).
__global__ void __launch_bounds__(128) k_front_small(DevCtx cx, const int4* __restrict__ tasks) {
    extern __shared__ double sm[];
    __shared__ double D[KMAX][KMAX + 1];
    int4 tk = tasks[blockIdx.x];
    const Front F = load_front(cx, tk.x);
    const int k = F.k, r = (int)F.r, f = (int)F.f;
    const int rp = ((r + 3) & ~3) + 4;
    double* Ls = sm;                  // Ls[c * rp + a] = L21[a][c]
    double* Ts = sm + KMAX * rp;      // Ts[c * rp + b] = U12[c][b]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        double a[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) a[c] = (lane < k && c < k) ? F.P[lane + (int64_t)c * f] : 0.0;
        warp_lu32(a, lane, k, F.c0, cx.flag);
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            D[lane][c] = a[c];
            if (lane < k && c < k) F.P[lane + (int64_t)c * f] = a[c];
        }
    } else {
        for (int c = warp - 1; c < k; c += 3) {
            const double* __restrict__ pl = F.P + k + (int64_t)c * f;
            const double* __restrict__ pt = F.T + (int64_t)c * r;
            for (int a = lane; a < r; a += 32) { Ls[c * rp + a] = pl[a]; Ts[c * rp + a] = pt[a]; }
        }
    }
    __syncthreads();
    for (int t = tid; t < 2 * r; t += 128) {   // row solves: first the r rows of L21, then of U12'
        double x[KMAX];
        const bool lower = t < r;
        const int a = lower ? t : t - r;
        double* row = (lower ? Ls : Ts) + a;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) x[c] = c < k ? row[c * rp] : 0.0;
        if (lower) trsm_row_upper(x, D, k); else trsm_row_lower_t(x, D, k);
        double* g = lower ? F.P + k + a : F.T + a;
        const int64_t ldg = lower ? f : r;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) if (c < k) { row[c * rp] = x[c]; g[(int64_t)c * ldg] = x[c]; }
    }
    __syncthreads();
    const int nt4 = (r + 3) >> 2;
    for (int t = tid; t < nt4 * nt4; t += 128) {
        const int a0 = (t % nt4) * 4, b0 = (t / nt4) * 4;
        double acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                acc[i][j] = (tk.y && a0 + i < r && b0 + j < r) ? F.C[(a0 + i) + (int64_t)(b0 + j) * r] : 0.0;
        for (int c = 0; c < k; ++c) {
            const double2 l01 = *reinterpret_cast<const double2*>(Ls + c * rp + a0);
            const double2 l23 = *reinterpret_cast<const double2*>(Ls + c * rp + a0 + 2);
            const double2 u01 = *reinterpret_cast<const double2*>(Ts + c * rp + b0);
            const double2 u23 = *reinterpret_cast<const double2*>(Ts + c * rp + b0 + 2);
            const double l[4] = {l01.x, l01.y, l23.x, l23.y}, u[4] = {u01.x, u01.y, u23.x, u23.y};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][j] -= l[i] * u[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (a0 + i < r && b0 + j < r) F.C[(a0 + i) + (int64_t)(b0 + j) * r] = acc[i][j];
    }
}

// ------------------------------------------------------------------ panel: pivot-block LU + TRSMs
// task: x = supernode, y = tile, z = number of L21 tiles (nt), w = total CTAs of this front.
// Tiles [0,nt) solve 128 rows of L21 against U11, tiles [nt,2nt) solve 128 rows of U12' against
// L11'.  Every CTA factors the (<=32x32) pivot block redundantly (warp 0, in registers) while the
// other warps already fetch their rows; the CTA that is last to have READ the unfactored block
// writes the factored one back (no CTA ever waits on another).
__global__ void __launch_bounds__(PANEL_ROWS) k_panel(DevCtx cx, const int4* __restrict__ tasks) {
    __shared__ double D[KMAX][KMAX + 1];
    __shared__ int s_last;
    int4 tk = tasks[blockIdx.x];
    const Front F = load_front(cx, tk.x);
    const int k = F.k, tid = threadIdx.x, lane = tid & 31;
    const int nt = tk.z;
    const bool lower = tk.y < nt;
    const int64_t row = (int64_t)(lower ? tk.y : tk.y - nt) * PANEL_ROWS + tid;
    double* src = lower ? F.P + F.k + row : F.T + row;
    const int64_t ld = lower ? F.f : F.r;
    double x[KMAX];
    if (tid < 32) {
        double a[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) a[c] = (lane < k && c < k) ? F.P[lane + (int64_t)c * F.f] : 0.0;
#pragma unroll
        for (int c = 0; c < KMAX; ++c) x[c] = (c < k && row < F.r) ? src[(int64_t)c * ld] : 0.0;
        warp_lu32(a, lane, k, F.c0, cx.flag);
#pragma unroll
        for (int c = 0; c < KMAX; ++c) D[lane][c] = a[c];
        if (lane == 0) {   // this CTA's reads of the unfactored block are complete
            __threadfence();
            int old = atomicAdd(cx.counters + tk.x, 1);
            s_last = ((old + 1) % tk.w) == 0;
        }
    } else {
#pragma unroll
        for (int c = 0; c < KMAX; ++c) x[c] = (c < k && row < F.r) ? src[(int64_t)c * ld] : 0.0;
    }
    __syncthreads();
    if (s_last)
        for (int e = tid; e < KMAX * k; e += PANEL_ROWS) { int i = e & 31, j = e >> 5; if (i < k) F.P[i + (int64_t)j * F.f] = D[i][j]; }
    if (row >= F.r) return;
    if (lower) trsm_row_upper(x, D, k); else trsm_row_lower_t(x, D, k);
#pragma unroll
    for (int c = 0; c < KMAX; ++c) if (c < k) src[(int64_t)c * ld] = x[c];
}

// ------------------------------------------------------------------ Schur update of the CB
// V (r x r) = beta*C - L21 (r x k) * U12 (k x r), U12 held transposed.  64x64 tile per CTA,
// 4x4 per thread, whole K (<= 32) staged in shared memory once.
// task: x = supernode, y = tile row, z = tile col, w = flags:
//   bit0 beta   : C holds assembled contributions (else it is taken as zero)
//   bit1 direct : V is written straight into the parent's front through the rel map (this front
//                 is its parent's only child, so nobody else writes there): panel entries +=,
//   bit2 assign : ... and the parent's contribution-block entries are assigned (=) instead of +=.
//   without bit1 V overwrites C in place.
__global__ void __launch_bounds__(256) k_gemm_cb(DevCtx cx, const int4* __restrict__ tasks) {
    __shared__ double As[KMAX][GEMM_TILE];
    __shared__ double Bs[KMAX][GEMM_TILE];
    int4 tk = tasks[blockIdx.x];
    const Front F = load_front(cx, tk.x);
    const int k = F.k, tid = threadIdx.x;
    const int64_t m0 = (int64_t)tk.y * GEMM_TILE, n0 = (int64_t)tk.z * GEMM_TILE;
    const double* __restrict__ A = F.P + F.k;
    const double* __restrict__ B = F.T;
    const int tx = tid & 15, ty = tid >> 4;
    const bool beta = tk.w & 1, direct = tk.w & 2, assign = tk.w & 4;
    double acc[4][4];
    // issue the C loads first so they are in flight while the operand tiles are staged
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t col = n0 + ty + 16 * j;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = m0 + tx + 16 * i;
            acc[i][j] = (beta && row < F.r && col < F.r) ? F.C[row + col * F.r] : 0.0;
        }
    }
    for (int e = tid; e < GEMM_TILE * k; e += 256) {
        int a = e & (GEMM_TILE - 1), p = e >> 6;
        int64_t ra = m0 + a, rb = n0 + a;
        As[p][a] = ra < F.r ? A[ra + (int64_t)p * F.f] : 0.0;
        Bs[p][a] = rb < F.r ? B[rb + (int64_t)p * F.r] : 0.0;
    }
    __syncthreads();
    for (int p = 0; p < k; ++p) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = As[p][tx + 16 * i]; b[i] = Bs[p][ty + 16 * i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] -= a[i] * b[j];
    }
    if (!direct) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + ty + 16 * j;
            if (col >= F.r) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t row = m0 + tx + 16 * i;
                if (row < F.r) F.C[row + col * F.r] = acc[i][j];
            }
        }
        return;
    }
    const Front Q = load_front(cx, cx.sn_parent[tk.x]);
    const int* __restrict__ rel = cx.rel + cx.rows_ptr[tk.x];
    int64_t prow[4], pcol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + tx + 16 * i, col = n0 + ty + 16 * i;
        prow[i] = row < F.r ? rel[row] : -1;
        pcol[i] = col < F.r ? rel[col] : -1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t pb = pcol[j];
        if (pb < 0) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t pa = prow[i];
            if (pa < 0) continue;
            const double v = acc[i][j];
            if (pb < Q.k) Q.P[pa + pb * Q.f] += v;
            else if (pa < Q.k) Q.T[(pb - Q.k) + pa * Q.r] += v;
            else {
                double* d = Q.C + (pa - Q.k) + (pb - Q.k) * Q.r;
                *d = assign ? v : *d + v;
            }
        }
    }
}

// ------------------------------------------------------------------ solves
__global__ void k_permute_scale(int n, const int* __restrict__ p, const double* __restrict__ Rs,
                                const double* __restrict__ b, double* __restrict__ w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { int pi = p[i]; w[i] = Rs[pi] * b[pi]; }
}
__global__ void k_unpermute(int n, const int* __restrict__ q, const double* __restrict__ w, double* __restrict__ x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[q[i]] = w[i];
}

// Forward substitution for one level.  task: x = supernode, y = row tile of the update vector.
// y_s = L11^{-1} (w[cols] + children contributions); upd_s = children contributions - L21 y_s.
// Every tile recomputes y_s (k <= 32); tile 0 stores it.  Children are gathered one at a time,
// ascending, so the summation order is fixed.  L11 is staged in shared memory up front so the
// 32-step substitution never waits on global memory.
__global__ void __launch_bounds__(FWD_ROWS) k_fwd(DevCtx cx, const int4* __restrict__ tasks,
                                                  const double* __restrict__ win, double* __restrict__ zout) {
    __shared__ double Ls[KMAX][KMAX + 1];
    __shared__ double ys[KMAX];
    __shared__ double acc[FWD_ROWS];
    int4 tk = tasks[blockIdx.x];
    const int s = tk.x;
    const Front F = load_front(cx, s);
    const int k = F.k, tid = threadIdx.x;
    const int64_t lo = (int64_t)tk.y * FWD_ROWS;            // first update row of this tile
    for (int e = tid; e < k * k; e += FWD_ROWS) { int i = e % k, j = e / k; Ls[i][j] = F.P[i + (int64_t)j * F.f]; }
    if (tid < KMAX) ys[tid] = tid < k ? win[F.c0 + tid] : 0.0;
    acc[tid] = 0.0;
    __syncthreads();
    for (int ci = cx.child_ptr[s]; ci < cx.child_ptr[s + 1]; ++ci) {
        const int c = cx.child_idx[ci];
        const int64_t rc = cx.rows_ptr[c + 1] - cx.rows_ptr[c];
        const int* __restrict__ rel = cx.rel + cx.rows_ptr[c];
        const double* __restrict__ uc = cx.upd + cx.rows_ptr[c];
        for (int64_t a = tid; a < rc; a += FWD_ROWS) {
            const int64_t ra = rel[a];
            if (ra < k) ys[ra] += uc[a];
            else if (ra - k >= lo && ra - k < lo + FWD_ROWS) acc[ra - k - lo] += uc[a];
        }
        __syncthreads();
    }
    if (tid < 32) {   // unit lower triangular solve with L11, one lane per row
        double y = tid < k ? ys[tid] : 0.0;
        for (int j = 0; j < k; ++j) {
            const double yj = __shfl_sync(0xffffffffu, y, j);
            if (tid > j && tid < k) y -= Ls[tid][j] * yj;
        }
        if (tid < k) {
            ys[tid] = y;
            if (tk.y == 0) zout[F.c0 + tid] = y;
        }
    }
    __syncthreads();
    const int64_t row = lo + tid;
    if (row < F.r) {
        double v = acc[tid];
        const double* __restrict__ src = F.P + F.k + row;
#pragma unroll 8
        for (int j = 0; j < k; ++j) v -= src[(int64_t)j * F.f] * ys[j];
        cx.upd[cx.rows_ptr[s] + row] = v;
    }
}

// Backward substitution for one level.  task: x = supernode, y = row tile, z = tiles of this
// supernode, w = slot of its partial sums in cx.bpart.   x[cols] = U11^{-1} (x[cols] - U12 x[rows]).
// Each CTA reduces BWD_ROWS rows of U12' against the gathered x; with several tiles the partial
// k-vectors go to scratch and the CTA that arrives last adds them in tile order (fixed summation
// order, nobody waits) and finishes the 32x32 back substitution.
__global__ void __launch_bounds__(BWD_ROWS) k_bwd(DevCtx cx, const int4* __restrict__ tasks, double* __restrict__ x) {
    __shared__ double Us[KMAX][KMAX + 1];
    __shared__ double xs[BWD_ROWS];
    __shared__ double part[KMAX];
    __shared__ int s_last;
    int4 tk = tasks[blockIdx.x];
    const int s = tk.x, ntiles = tk.z;
    const Front F = load_front(cx, s);
    const int k = F.k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t lo = (int64_t)tk.y * BWD_ROWS;
    const int cnt = (int)(F.r - lo < BWD_ROWS ? F.r - lo : BWD_ROWS);
    const int* __restrict__ rows = cx.rows + cx.rows_ptr[s] + lo;
    if (tid < cnt) xs[tid] = x[rows[tid]];
    for (int e = tid; e < k * k; e += BWD_ROWS) { int i = e % k, j = e / k; Us[i][j] = F.P[i + (int64_t)j * F.f]; }
    __syncthreads();
    for (int i = warp; i < k; i += BWD_ROWS / 32) {
        const double* __restrict__ col = F.T + (int64_t)i * F.r + lo;
        double v = 0.0;
        for (int a = lane; a < cnt; a += 32) v += col[a] * xs[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) part[i] = v;
    }
    __syncthreads();
    if (ntiles > 1) {
        double* slot = cx.bpart + (int64_t)tk.w * KMAX;
        if (tid < k) slot[(int64_t)tk.y * KMAX + tid] = part[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int old = atomicAdd(cx.counters2 + s, 1);
            s_last = ((old + 1) % ntiles) == 0;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (tid < k) {
            double v = 0.0;
            for (int t = 0; t < ntiles; ++t) v += __ldcg(slot + (int64_t)t * KMAX + tid);
            part[tid] = v;
        }
        __syncthreads();
    }
    if (tid < 32) {
        double v = tid < k ? x[F.c0 + tid] - part[tid] : 0.0;
        for (int j = k - 1; j >= 0; --j) {
            double xj = 0.0;
            if (tid == j) xj = v / Us[j][j];
            xj = __shfl_sync(0xffffffffu, xj, j);
            if (tid == j) v = xj;
            if (tid < j) v -= Us[tid][j] * xj;
        }
        if (tid < k) x[F.c0 + tid] = v;
    }
}

constexpr int SMALL_F_MAX = 96;

}  // namespace

int front_small_limit() { return SMALL_F_MAX; }

static size_t small_smem(int rmax) { return sizeof(double) * 2 * KMAX * (size_t)(((rmax + 3) & ~3) + 4); }

cudaError_t kernels_init() {
    return cudaFuncSetAttribute(k_front_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem(SMALL_F_MAX));
}

void launch_rowscale(cudaStream_t st, int n, const int64_t* rowptr, const int64_t* rowidx, const double* av, double* Rs) {
    k_rowscale<<<(n + 255) / 256, 256, 0, st>>>(n, rowptr, rowidx, av, Rs);
}
void launch_scatter(cudaStream_t st, int64_t nnz, const int64_t* dst, const int* arow, const double* Rs,
                    const double* av, double* lu) {
    int64_t blocks = (nnz + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_scatter<<<(int)blocks, 256, 0, st>>>(nnz, dst, arow, Rs, av, lu);
}
void launch_zero_cb(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) k_zero_cb<<<ntasks, 256, 0, st>>>(cx, tasks);
}
void launch_extend_add(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) k_extend_add<<<ntasks, 256, 0, st>>>(cx, tasks);
}
void launch_front_small(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, int fmax) {
    if (ntasks <= 0) return;
    k_front_small<<<ntasks, 128, small_smem(fmax), st>>>(cx, tasks);   // fmax bounds r of the class
}
void launch_panel(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) k_panel<<<ntasks, PANEL_ROWS, 0, st>>>(cx, tasks);
}
void launch_gemm_cb(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks) {
    if (ntasks > 0) k_gemm_cb<<<ntasks, 256, 0, st>>>(cx, tasks);
}
void launch_permute_scale(cudaStream_t st, int n, const int* p, const double* Rs, const double* b, double* w) {
    k_permute_scale<<<(n + 255) / 256, 256, 0, st>>>(n, p, Rs, b, w);
}
void launch_unpermute(cudaStream_t st, int n, const int* q, const double* w, double* x) {
    k_unpermute<<<(n + 255) / 256, 256, 0, st>>>(n, q, w, x);
}
void launch_fwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, const double* win, double* zout) {
    if (ntasks > 0) k_fwd<<<ntasks, FWD_ROWS, 0, st>>>(cx, tasks, win, zout);
}
void launch_bwd(cudaStream_t st, const DevCtx& cx, const int4* tasks, int ntasks, double* x) {
    if (ntasks > 0) k_bwd<<<ntasks, BWD_ROWS, 0, st>>>(cx, tasks, x);
}

}  // namespace smslu
