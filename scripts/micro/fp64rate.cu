// FP64 pipe microbenchmark: DFMA (CUDA cores) vs DMMA (mma.sync m8n8k4 f64) throughput, and latency.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_tp(double* out, int iters) {
    double a[8], b = 1.0000001, c = 1e-9;
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_lat(double* out, long long* cyc, int iters) {
    double a = threadIdx.x, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) a = fma(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = (t1 - t0);
}
__global__ void ddiv_lat(double* out, long long* cyc, int iters) {
    double a = threadIdx.x + 3.0, b = 1.0000001;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) a = b / a + 2.0;
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = (t1 - t0);
}
__global__ void dmma_tp(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_lat(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6, c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1; if (threadIdx.x == 0) *cyc = (t1 - t0);
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 8 * 1024 * 8); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms; long long c;
    for (int bps = 1; bps <= 8; bps *= 2) {
        int iters = 20000, grid = 148 * bps, thr = 256;
        dfma_tp<<<grid, thr>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); dfma_tp<<<grid, thr>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  %d CTA/SM x256thr: %.2f TFLOP/s\n", bps, 2.0 * grid * thr * 8.0 * iters / (ms * 1e-3) / 1e12);
        dmma_tp<<<grid, thr>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); dmma_tp<<<grid, thr>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  %d CTA/SM x256thr: %.2f TFLOP/s\n", bps, 2.0 * grid * (thr / 32) * 8.0 * 256.0 * iters / (ms * 1e-3) / 1e12);
    }
    dfma_lat<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency %.1f cycles\n", c / 10000.0);
    ddiv_lat<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("DDIV+DADD dependent latency %.1f cycles\n", c / 10000.0);
    dmma_lat<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("DMMA dependent latency %.1f cycles\n", c / 10000.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
