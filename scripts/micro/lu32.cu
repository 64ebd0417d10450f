// Microbenchmark: ways to factor a 32x32 pivot block and to solve 128 rows against it (one CTA).
#include <cstdio>
#include <cuda_runtime.h>
#define K 32
__device__ __forceinline__ void warp_lu32(double (&a)[K], int lane) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double piv = __shfl_sync(0xffffffffu, a[j], j);
        double l = 0.0;
        if (lane > j) { l = a[j] / piv; a[j] = l; }
#pragma unroll
        for (int c = j + 1; c < K; ++c) { const double u = __shfl_sync(0xffffffffu, a[c], j); if (lane > j) a[c] -= l * u; }
    }
}
__global__ void v1(const double* A, double* out, long long* cyc, int reps) {
    __shared__ double D[K][K + 1];
    int lane = threadIdx.x;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (threadIdx.x < 32) {
            double a[K];
#pragma unroll
            for (int c = 0; c < K; ++c) a[c] = A[lane + c * K] + r * 1e-9;
            warp_lu32(a, lane);
#pragma unroll
            for (int c = 0; c < K; ++c) D[lane][c] = a[c];
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = (t1 - t0) / reps;
    if (threadIdx.x < 32) for (int c = 0; c < K; ++c) out[lane + c * K] = D[lane][c];
}
// single warp, rows in shared memory, rolled loops
__global__ void v2(const double* A, double* out, long long* cyc, int reps, int use_rcp) {
    __shared__ double D[K][K + 1];
    int lane = threadIdx.x;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (threadIdx.x < 32) {
            for (int c = 0; c < K; ++c) D[lane][c] = A[lane + c * K] + r * 1e-9;
            __syncwarp();
            for (int j = 0; j < K; ++j) {
                const double piv = D[j][j];
                if (lane > j) {
                    const double l = use_rcp ? D[lane][j] * (1.0 / piv) : D[lane][j] / piv;
                    D[lane][j] = l;
#pragma unroll 4
                    for (int c = j + 1; c < K; ++c) D[lane][c] -= l * D[j][c];
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = (t1 - t0) / reps;
    if (threadIdx.x < 32) for (int c = 0; c < K; ++c) out[lane + c * K] = D[lane][c];
}
// 128 threads, 2 barriers per step (first version of the panel kernel)
__global__ void v3(const double* A, double* out, long long* cyc, int reps) {
    __shared__ double D[K][K + 1];
    int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int e = tid; e < K * K; e += 128) D[e % K][e / K] = A[e] + r * 1e-9;
        __syncthreads();
        for (int j = 0; j < K; ++j) {
            const double piv = D[j][j];
            if (tid > j && tid < K) D[tid][j] /= piv;
            __syncthreads();
            for (int c = j + 1 + ty; c < K; c += 4) { int i = j + 1 + tx; if (i < K) D[i][c] -= D[i][j] * D[j][c]; }
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) *cyc = (t1 - t0) / reps;
    if (tid < 32) for (int c = 0; c < K; ++c) out[tid + c * K] = D[tid][c];
}
// TRSM variants: 128 rows x 32 against upper factor in D
__global__ void t1(const double* A, double* X, long long* cyc, int reps) {
    __shared__ double D[K][K + 1];
    int tid = threadIdx.x;
    for (int e = tid; e < K * K; e += 128) D[e % K][e / K] = A[e] + (e % K == e / K ? 8.0 : 0.0);
    __syncthreads();
    long long t0 = clock64();
    double x[K];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < K; ++c) x[c] = X[tid + c * 128] + r;
#pragma unroll
        for (int c = 0; c < K; ++c) { double v = x[c];
#pragma unroll
            for (int p = 0; p < c; ++p) v -= x[p] * D[p][c];
            x[c] = v / D[c][c]; }
#pragma unroll
        for (int c = 0; c < K; ++c) X[tid + c * 128] = x[c];
    }
    long long t1_ = clock64();
    if (tid == 0) *cyc = (t1_ - t0) / reps;
}
__global__ void t2(const double* A, double* X, long long* cyc, int reps) {   // rolled, x in smem, reciprocal diag
    __shared__ double D[K][K + 1];
    __shared__ double xs[K][128];
    __shared__ double rd[K];
    int tid = threadIdx.x;
    for (int e = tid; e < K * K; e += 128) D[e % K][e / K] = A[e] + (e % K == e / K ? 8.0 : 0.0);
    __syncthreads();
    if (tid < K) rd[tid] = 1.0 / D[tid][tid];
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int c = 0; c < K; ++c) xs[c][tid] = X[tid + c * 128] + r;
        for (int c = 0; c < K; ++c) { double v = xs[c][tid];
#pragma unroll 4
            for (int p = 0; p < c; ++p) v -= xs[p][tid] * D[p][c];
            xs[c][tid] = v * rd[c]; }
        for (int c = 0; c < K; ++c) X[tid + c * 128] = xs[c][tid];
    }
    long long t1_ = clock64();
    if (tid == 0) *cyc = (t1_ - t0) / reps;
}
// TRSM as a product with the explicit inverse (no dependent chain): x_new[c] = sum_p x[p] * Uinv[p][c]
__global__ void t3(const double* A, double* X, long long* cyc, int reps) {
    __shared__ double D[K][K + 1];
    int tid = threadIdx.x;
    for (int e = tid; e < K * K; e += 128) D[e % K][e / K] = (e % K <= e / K) ? A[e] : 0.0;
    __syncthreads();
    long long t0 = clock64();
    double x[K], y[K];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < K; ++c) x[c] = X[tid + c * 128] + r;
#pragma unroll
        for (int c = 0; c < K; ++c) { double v = 0;
#pragma unroll
            for (int p = 0; p <= c; ++p) v += x[p] * D[p][c];
            y[c] = v; }
#pragma unroll
        for (int c = 0; c < K; ++c) X[tid + c * 128] = y[c];
    }
    long long t1_ = clock64();
    if (tid == 0) *cyc = (t1_ - t0) / reps;
}
int main() {
    double *A, *out, *X; long long* cyc;
    cudaMalloc(&A, K * K * 8); cudaMalloc(&out, K * K * 8); cudaMalloc(&X, 128 * K * 8); cudaMalloc(&cyc, 8);
    double h[K * K]; for (int i = 0; i < K * K; ++i) h[i] = ((i * 7919) % 101) / 101.0 + ((i % K == i / K) ? 40.0 : 0.0);
    cudaMemcpy(A, h, sizeof h, cudaMemcpyHostToDevice); cudaMemset(X, 0, 128 * K * 8);
    long long c; int reps = 200;
    auto rep = [&](const char* n) { cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %8lld cycles  (%s)\n", n, c, cudaGetErrorString(cudaGetLastError())); };
    for (int w = 0; w < 2; ++w) {
        v1<<<1, 128>>>(A, out, cyc, reps); rep("LU v1 warp registers+shuffles");
        v2<<<1, 128>>>(A, out, cyc, reps, 0); rep("LU v2 one warp smem rolled, div");
        v2<<<1, 128>>>(A, out, cyc, reps, 1); rep("LU v2 one warp smem rolled, rcp");
        v3<<<1, 128>>>(A, out, cyc, reps); rep("LU v3 128 threads 2 barriers");
        t1<<<1, 128>>>(A, X, cyc, reps); rep("TRSM t1 registers unrolled div");
        t2<<<1, 128>>>(A, X, cyc, reps); rep("TRSM t2 smem rolled rcp");
        t3<<<1, 128>>>(A, X, cyc, reps); rep("TRSM t3 inverse product unrolled");
        v1<<<1, 128>>>(A, out, cyc, 1); rep("LU v1 single cold rep");
        t1<<<1, 128>>>(A, X, cyc, 1); rep("TRSM t1 single cold rep");
        t2<<<1, 128>>>(A, X, cyc, 1); rep("TRSM t2 single cold rep");
    }
    return 0;
}
