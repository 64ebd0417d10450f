// Microbenchmark of the pieces of one panel step (one CTA of 128 threads): candidates for k_panel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_parts panel_parts.cu
#include <cstdio>
#include <cuda_runtime.h>
#define K 32
#define CLD 34

__device__ __forceinline__ double rcp(double x) { return 1.0 / x; }

// LU-A: lane = row, warp = strip of 8 columns held in registers; the multipliers of the current
// step travel through shared memory (double buffered), the pivot row by shuffles; one barrier/step.
__device__ __forceinline__ void lu_strips(double (&a)[8], double (*colbuf)[K], double* rd, int lane, int warp) {
#pragma unroll 1
    for (int jw = 0; jw < 4; ++jw) {
#pragma unroll
        for (int uj = 0; uj < 8; ++uj) {
            const int j = jw * 8 + uj;
            if (warp == jw) {
                const double piv = __shfl_sync(0xffffffffu, a[uj], j);
                const double rinv = rcp(piv);
                const double l = lane > j ? a[uj] * rinv : 0.0;
                if (lane > j) a[uj] = l;
                colbuf[j & 1][lane] = l;
                if (lane == j) rd[j] = rinv;
            }
            __syncthreads();
            const double l = colbuf[j & 1][lane];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double uv = __shfl_sync(0xffffffffu, a[u], j);
                if (warp * 8 + u > j) a[u] -= l * uv;       // l == 0 for lane <= j
            }
        }
    }
}
__global__ void luA(const double* A, double* out, long long* cyc, int reps) {
    __shared__ double colbuf[2][K];
    __shared__ double rd[K];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tot = 0;
    double a[8];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = A[lane + (warp * 8 + u) * K] + r * 1e-9;
        __syncthreads();
        long long t0 = clock64();
        lu_strips(a, colbuf, rd, lane, warp);
        __syncthreads();
        tot += clock64() - t0;
    }
    if (tid == 0) *cyc = tot / reps;
#pragma unroll
    for (int u = 0; u < 8; ++u) out[lane + (warp * 8 + u) * K] = a[u];
}
// LU-B: one warp, lane = row, all 32 columns in registers, fully unrolled, shuffles only.
__global__ void luB(const double* A, double* out, long long* cyc, int reps) {
    const int tid = threadIdx.x, lane = tid & 31;
    long long tot = 0;
    double a[K];
    if (tid >= 32) return;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < K; ++c) a[c] = A[lane + c * K] + r * 1e-9;
        __syncwarp();
        long long t0 = clock64();
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double piv = __shfl_sync(0xffffffffu, a[j], j);
            const double l = lane > j ? a[j] * rcp(piv) : 0.0;
            if (lane > j) a[j] = l;
#pragma unroll
            for (int c = j + 1; c < K; ++c) a[c] -= l * __shfl_sync(0xffffffffu, a[c], j);
        }
        tot += clock64() - t0;
    }
    if (tid == 0) *cyc = tot / reps;
#pragma unroll
    for (int c = 0; c < K; ++c) out[lane + c * K] = a[c];
}
// LU-B2/B3: as LU-B, but the reciprocal of the next pivot is started as soon as column j+1 is up to date,
// so that its latency overlaps the rest of the step (B3: __drcp_rn instead of 1.0/x).
template <int RCPK>
__global__ void luB2(const double* A, double* out, long long* cyc, int reps) {
    const int tid = threadIdx.x, lane = tid & 31;
    long long tot = 0;
    double a[K];
    if (tid >= 32) return;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < K; ++c) a[c] = A[lane + c * K] + r * 1e-9;
        __syncwarp();
        long long t0 = clock64();
        double piv = __shfl_sync(0xffffffffu, a[0], 0);
        double rinv = RCPK ? __drcp_rn(piv) : 1.0 / piv;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double l = lane > j ? a[j] * rinv : 0.0;
            if (lane > j) a[j] = l;
            if (j + 1 < K) {
                a[j + 1] -= l * __shfl_sync(0xffffffffu, a[j + 1], j);
                piv = __shfl_sync(0xffffffffu, a[j + 1], j + 1);
                rinv = RCPK ? __drcp_rn(piv) : 1.0 / piv;
            }
#pragma unroll
            for (int c = j + 2; c < K; ++c) a[c] -= l * __shfl_sync(0xffffffffu, a[c], j);
        }
        tot += clock64() - t0;
    }
    if (tid == 0) *cyc = tot / reps;
#pragma unroll
    for (int c = 0; c < K; ++c) out[lane + c * K] = a[c];
}
// LU-D: lane = COLUMN, all 32 rows in registers: the multipliers of step j live in lane j and are broadcast
// by shuffles, the pivot-row element is the lane's own register.
__global__ void luD(const double* A, double* out, long long* cyc, int reps) {
    const int tid = threadIdx.x, lane = tid & 31;
    long long tot = 0;
    double a[K];      // a[i] = D[i][lane]
    if (tid >= 32) return;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < K; ++i) a[i] = A[i + lane * K] + r * 1e-9;
        __syncwarp();
        long long t0 = clock64();
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double rinv = 1.0 / a[j];                 // only lane j's value is used
            const double u = lane > j ? a[j] : 0.0;         // pivot row element of this column
#pragma unroll
            for (int i = j + 1; i < K; ++i) {
                const double l = __shfl_sync(0xffffffffu, a[i] * rinv, j);
                if (lane == j) a[i] = l;
                a[i] -= l * u;
            }
        }
        tot += clock64() - t0;
    }
    if (tid == 0) *cyc = tot / reps;
#pragma unroll
    for (int i = 0; i < K; ++i) out[i + lane * K] = a[i];
}
// LU-K: the pivot-warp code of k_panel verbatim (row from shared memory, identity padding, branch-free bad-pivot
// bookkeeping, factor rows written back to shared memory), warp 0 of a 256-thread CTA.
__device__ __forceinline__ bool bad_pivot(double p) { return !(fabs(p) > 0.0) || !isfinite(p); }
__global__ void luK(const double* A, double* out, long long* cyc, int reps, int kind, int w) {
    __shared__ __align__(16) double D[K][CLD];
    __shared__ __align__(16) double W[K][CLD];
    __shared__ double rd[K];
    __shared__ int flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tot = 0;
    double x[K];
    for (int r = 0; r < reps; ++r) {
        for (int e = tid; e < K * K; e += blockDim.x) D[e % K][e / K] = A[e] + r * 1e-9;
        __syncthreads();
        long long t0 = clock64();
        if (warp == 0) {
#pragma unroll
            for (int c2 = 0; c2 < K / 2; ++c2) {
                const double2 v = *reinterpret_cast<const double2*>(&D[lane][2 * c2]);
                x[2 * c2] = v.x; x[2 * c2 + 1] = v.y;
            }
#pragma unroll
            for (int c = 0; c < K; ++c) x[c] = (lane >= w && c == lane) ? 1.0 : x[c];
            double myr = 0.0;
            int bad = K;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double piv = __shfl_sync(0xffffffffu, x[j], j);
                const double rinv = 1.0 / piv;
                bad = (bad == K && bad_pivot(piv)) ? j : bad;
                myr = lane == j ? rinv : myr;
                const double l = lane > j ? x[j] * rinv : 0.0;
                x[j] = lane > j ? l : x[j];
#pragma unroll
                for (int c = j + 1; c < K; ++c) x[c] -= l * __shfl_sync(0xffffffffu, x[c], j);
            }
            rd[lane] = myr;
            if (lane == 0 && bad < w) atomicMin(&flag, bad);
            if (kind == 0) {
#pragma unroll
                for (int c2 = 0; c2 < K / 2; ++c2) *reinterpret_cast<double2*>(&W[lane][2 * c2]) = make_double2(x[2 * c2], x[2 * c2 + 1]);
            } else {
#pragma unroll
                for (int p = 0; p < K; ++p) W[p][lane] = x[p];
            }
        }
        tot += clock64() - t0;
        __syncthreads();
    }
    if (tid == 0) *cyc = tot / reps;
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < K; ++c) out[lane + c * K] = x[c];
    }
    if (tid == 300) out[0] = W[1][1] + rd[1];
}
// LU-C: as LU-A but the whole block lives in shared memory D[i][CLD]; thread (lane = row, warp = 8 columns)
// reads the pivot row with 128-bit loads; one barrier per step, reciprocal produced one step ahead.
__global__ void luC(const double* A, double* out, long long* cyc, int reps) {
    __shared__ __align__(16) double D[K][CLD];
    __shared__ double rd[K];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tot = 0;
    for (int r = 0; r < reps; ++r) {
        for (int e = tid; e < K * K; e += 128) D[e % K][e / K] = A[e] + r * 1e-9;
        __syncthreads();
        long long t0 = clock64();
        if (tid == 0) rd[0] = rcp(D[0][0]);
        __syncthreads();
        double a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = D[lane][warp * 8 + u];
        for (int j = 0; j < K; ++j) {
            const double l = lane > j ? D[lane][j] * rd[j] : 0.0;    // column j is final since step j-1
            const double2* pr = reinterpret_cast<const double2*>(&D[j][warp * 8]);
            const double2 u01 = pr[0], u23 = pr[1], u45 = pr[2], u67 = pr[3];
            const double uu[8] = {u01.x, u01.y, u23.x, u23.y, u45.x, u45.y, u67.x, u67.y};
#pragma unroll
            for (int u = 0; u < 8; ++u) if (warp * 8 + u > j) a[u] -= l * uu[u];
            // publish what the next step reads: column j+1 (all rows) and row j+1 (all columns)
            const int jn = j + 1;
            if (jn < K) {
                if (warp == (jn >> 3)) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (u == (jn & 7)) { D[lane][jn] = a[u]; if (lane == jn) rd[jn] = rcp(a[u]); }
                }
                if (lane == jn) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) D[jn][warp * 8 + u] = a[u];
                }
            }
            __syncthreads();
        }
        tot += clock64() - t0;
#pragma unroll
        for (int u = 0; u < 8; ++u) D[lane][warp * 8 + u] = (lane > warp * 8 + u) ? D[lane][warp * 8 + u] * rd[warp * 8 + u] : a[u];
        __syncthreads();
    }
    if (tid == 0) *cyc = tot / reps;
    for (int e = tid; e < K * K; e += 128) out[e] = D[e % K][e / K];
}
// TRSM-T: thread = row, 32 values in registers, factors W[p][c] (stride 34) read with 128-bit loads.
__global__ void trsmT(const double* A, double* X, long long* cyc, int reps) {
    __shared__ __align__(16) double W[K][CLD];
    __shared__ double rd[K];
    const int tid = threadIdx.x;
    for (int e = tid; e < K * K; e += 128) W[e % K][e / K] = A[e];
    __syncthreads();
    if (tid < K) rd[tid] = 1.0 / W[tid][tid];
    __syncthreads();
    long long tot = 0;
    double x[K];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < K; ++c) x[c] = X[tid + c * 128] + r;
        __syncthreads();
        long long t0 = clock64();
#pragma unroll
        for (int p = 0; p < K; ++p) {
            const double xp = x[p] * rd[p];
            x[p] = xp;
            const double2* wr = reinterpret_cast<const double2*>(&W[p][0]);
#pragma unroll
            for (int c2 = (p + 1) / 2; c2 < K / 2; ++c2) {
                const double2 w = wr[c2];
                if (2 * c2 > p) x[2 * c2] -= xp * w.x;
                x[2 * c2 + 1] -= xp * w.y;
            }
        }
        tot += clock64() - t0;
#pragma unroll
        for (int c = 0; c < K; ++c) X[tid + c * 128] = x[c];
    }
    if (tid == 0) *cyc = tot / reps;
}
// UPD-U1: left-looking update of 128 rows x 32 columns with j0 = 96 earlier columns, thread = row,
// coefficients cf[m][CLD] read with 128-bit loads, row values streamed from global (L2-resident).
__global__ void updU1(const double* Rg, double* X, long long* cyc, int reps, int j0) {
    extern __shared__ __align__(16) double cf[];
    const int tid = threadIdx.x;
    for (int e = tid; e < j0 * CLD; e += 128) cf[e] = 1e-3 * ((e * 31) % 17);
    __syncthreads();
    long long tot = 0;
    double x[K];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < K; ++c) x[c] = r;
        __syncthreads();
        long long t0 = clock64();
        const double* base = Rg + tid;
        for (int m0 = 0; m0 < j0; m0 += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = base[(m0 + u) * 2048];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double2* cm = reinterpret_cast<const double2*>(cf + (m0 + u) * CLD);
#pragma unroll
                for (int c2 = 0; c2 < K / 2; ++c2) { const double2 w = cm[c2]; x[2 * c2] -= v[u] * w.x; x[2 * c2 + 1] -= v[u] * w.y; }
            }
        }
        tot += clock64() - t0;
#pragma unroll
        for (int c = 0; c < K; ++c) X[tid + c * 128] = x[c];
    }
    if (tid == 0) *cyc = tot / reps;
}
// UPD-U2: same update on the FP64 tensor pipe: warp = 32 rows x 32 columns = 4 x 4 DMMA tiles.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void updU2(const double* Rg, double* X, long long* cyc, int reps, int j0) {
    extern __shared__ __align__(16) double cf[];      // cf[m][CLD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fc = lane & 3;
    for (int e = tid; e < j0 * CLD; e += 128) cf[e] = 1e-3 * ((e * 31) % 17);
    __syncthreads();
    long long tot = 0;
    double acc[4][4][2];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = r;
        __syncthreads();
        long long t0 = clock64();
        const double* base = Rg + warp * 32 + fr;
        for (int m0 = 0; m0 < j0; m0 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = -base[8 * i + (m0 + fc) * 2048];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = cf[(m0 + fc) * CLD + 8 * j + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        tot += clock64() - t0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { X[(warp * 32 + 8 * i + fr) + (8 * j + 2 * fc) * 128] = acc[i][j][0]; X[(warp * 32 + 8 * i + fr) + (8 * j + 2 * fc + 1) * 128] = acc[i][j][1]; }
    }
    if (tid == 0) *cyc = tot / reps;
}
int main() {
    double *A, *out, *X, *Rg; long long* cyc;
    cudaMalloc(&A, K * K * 8); cudaMalloc(&out, K * K * 8); cudaMalloc(&X, 128 * K * 8); cudaMalloc(&cyc, 8);
    cudaMalloc(&Rg, 2048 * 128 * 8); cudaMemset(Rg, 0, 2048 * 128 * 8);
    double h[K * K]; for (int i = 0; i < K * K; ++i) h[i] = ((i * 7919) % 101) / 101.0 + ((i % K == i / K) ? 40.0 : 0.0);
    cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice); cudaMemset(X, 0, 128 * K * 8);
    double ref[K * K], got[K * K];
    for (int i = 0; i < K * K; ++i) ref[i] = h[i];        // column-major reference LU
    for (int j = 0; j < K; ++j) for (int i = j + 1; i < K; ++i) { ref[i + j * K] /= ref[j + j * K]; for (int c = j + 1; c < K; ++c) ref[i + c * K] -= ref[i + j * K] * ref[j + c * K]; }
    long long c; int reps = 100;
    auto rep = [&](const char* n, bool chk) {
        cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double err = 0; if (chk) { cudaMemcpy(got, out, sizeof got, cudaMemcpyDeviceToHost); for (int i = 0; i < K * K; ++i) { double d = got[i] - ref[i]; if (d < 0) d = -d; if (d > err) err = d; } }
        printf("%-48s %8lld cycles  err %.1e (%s)\n", n, c, err, cudaGetErrorString(cudaGetLastError())); };
    for (int w = 0; w < 2; ++w) {
        luA<<<1, 128>>>(A, out, cyc, reps); rep("LU-A 4 warps, register strips, 1 barrier/step", true);
        luB<<<1, 128>>>(A, out, cyc, reps); rep("LU-B 1 warp, registers + shuffles, unrolled", true);
        luC<<<1, 128>>>(A, out, cyc, reps); rep("LU-C 4 warps, smem block + register strips", true);
        luB2<0><<<1, 128>>>(A, out, cyc, reps); rep("LU-B2 1 warp, early reciprocal (1.0/x)", true);
        luB2<1><<<1, 128>>>(A, out, cyc, reps); rep("LU-B3 1 warp, early reciprocal (__drcp_rn)", true);
        luD<<<1, 128>>>(A, out, cyc, reps); rep("LU-D 1 warp, lane = column", true);
        luK<<<1, 256>>>(A, out, cyc, reps, 0, 32); rep("LU-K k_panel pivot warp verbatim, hot (100 reps)", true);
        luK<<<1, 256>>>(A, out, cyc, 1, 0, 32); rep("LU-K k_panel pivot warp verbatim, single cold rep", false);
        luB<<<1, 128>>>(A, out, cyc, 1); rep("LU-B single cold rep", false);
        trsmT<<<1, 128>>>(A, X, cyc, reps); rep("TRSM-T registers, 128-bit factor loads", false);
        updU1<<<1, 128, 96 * CLD * 8>>>(Rg, X, cyc, reps, 96); rep("UPD-U1 j0=96 thread=row, 128-bit coef loads", false);
        updU2<<<1, 128, 96 * CLD * 8>>>(Rg, X, cyc, reps, 96); rep("UPD-U2 j0=96 DMMA", false);
    }
    return 0;
}
