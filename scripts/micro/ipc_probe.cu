// Two processes, one GPU each: does CUDA IPC peer mapping work on this box, how fast are remote stores over
// NVLink, and what does a cross-GPU flag round trip cost?  (Design input for the peer-store replication of the
// top-of-tree panels; see DESIGN.md "Multi-GPU".)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ipc_probe ipc_probe.cu ; run: ./ipc_probe
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "rank %d: %s -> %s\n", rank, #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__global__ void k_fill_remote(double* remote, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) remote[i] = v + (double)i;
}
__global__ void k_signal(volatile int* remote_flag, int v) {
    __threadfence_system();
    *remote_flag = v;
}
__global__ void k_wait(volatile int* flag, int v) {
    while (*flag < v) { }
}
__global__ void k_check(const double* buf, size_t n, double v, int* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) if (buf[i] != v + (double)i) atomicAdd(bad, 1);
}
// ping-pong inside one kernel per side: rank 0 writes k to the peer and waits for k to come back
__global__ void k_pingpong(volatile int* mine, volatile int* remote, int rank, int iters) {
    for (int k = 1; k <= iters; ++k) {
        if (rank == 0) { *remote = k; __threadfence_system(); while (*mine < k) { } }
        else { while (*mine < k) { } *remote = k; __threadfence_system(); }
    }
}

int main() {
    int p01[2], p10[2];
    if (pipe(p01) || pipe(p10)) return 1;
    pid_t child = fork();
    const int rank = child == 0 ? 1 : 0;
    const int rd = rank == 0 ? p10[0] : p01[0], wr = rank == 0 ? p01[1] : p10[1];
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { if (rank == 0) printf("ipc_probe: needs 2 GPUs (have %d)\n", ndev); return 0; }
    CK(cudaSetDevice(rank));
    const size_t n = (size_t)32 << 20;     // 256 MB of doubles
    double* buf; int* flags; int* bad;
    CK(cudaMalloc(&buf, n * sizeof(double)));
    CK(cudaMalloc(&flags, 64 * sizeof(int)));
    CK(cudaMalloc(&bad, sizeof(int)));
    CK(cudaMemset(flags, 0, 64 * sizeof(int)));
    CK(cudaMemset(bad, 0, sizeof(int)));
    CK(cudaMemset(buf, 0, n * sizeof(double)));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hb, hf, pb, pf;
    CK(cudaIpcGetMemHandle(&hb, buf));
    CK(cudaIpcGetMemHandle(&hf, flags));
    if (write(wr, &hb, sizeof hb) != sizeof hb || write(wr, &hf, sizeof hf) != sizeof hf) return 3;
    if (read(rd, &pb, sizeof pb) != sizeof pb || read(rd, &pf, sizeof pf) != sizeof pf) return 3;
    double* rbuf; int* rflags;
    CK(cudaIpcOpenMemHandle((void**)&rbuf, pb, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle((void**)&rflags, pf, cudaIpcMemLazyEnablePeerAccess));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // 1. remote stores: each rank fills the peer's buffer, signals, waits for its own to be filled, checks
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_fill_remote<<<148 * 8, 256>>>(rbuf, n, 100.0 * (rep + 1) + rank);
        k_signal<<<1, 1>>>(rflags, rep + 1);
        CK(cudaEventRecord(e1));
        k_wait<<<1, 1>>>(flags, rep + 1);
        k_check<<<148 * 8, 256>>>(buf, n, 100.0 * (rep + 1) + (1 - rank), bad);
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        int hbad; CK(cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost));
        printf("rank %d rep %d: remote store of %zu MB in %.3f ms = %.1f GB/s, mismatches %d\n", rank, rep, n * 8 >> 20, ms, n * 8 / ms / 1e6, hbad);
        fflush(stdout);
        // both sides must be done checking before the next rep overwrites: handshake through the pipes
        char c = 1; if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 3;
    }
    // 2. flag round trip
    CK(cudaMemset(flags + 32, 0, sizeof(int)));
    CK(cudaDeviceSynchronize());
    { char c = 1; if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 3; }
    const int iters = 2000;
    CK(cudaEventRecord(e0));
    k_pingpong<<<1, 1>>>(flags + 32, rflags + 32, rank, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("rank %d: %d flag round trips in %.3f ms = %.2f us per round trip\n", rank, iters, ms, 1e3 * ms / iters);
    { char c = 1; if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 3; }
    CK(cudaIpcCloseMemHandle(rbuf)); CK(cudaIpcCloseMemHandle(rflags));
    if (rank == 0) { int st; waitpid(child, &st, 0); printf("ipc_probe: child exit %d\n", WEXITSTATUS(st)); }
    return 0;
}
