// l2_window.cu -- does data written by one phase survive in the 126 MB L2 while X MB of other traffic streams through?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2_window l2_window.cu && ./l2_window
// One persistent kernel (296 blocks, software grid barrier): phase 1 writes A (S MB), phase 2 streams X MB (half read, half
// written) through B, phase 3 reads A back and is timed with %globaltimer.  The time of phase 3 against X shows the
// residency window that a fused H-pass -> E-pass wavefront can count on.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ void grid_bar(unsigned *bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (*((volatile unsigned *)bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256) win(double2 *A, size_t nA, double2 *B, size_t nS, unsigned *bar, unsigned long long *out, int rd_only) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    if (!rd_only) for (size_t i = tid; i < nA; i += nt) A[i] = make_double2((double)i, 1.0);
    grid_bar(bar, gridDim.x);
    for (size_t i = tid; i < nS; i += nt) { double2 v = B[i]; v.x += 1.0; B[i + nS] = v; }
    grid_bar(bar, 2 * gridDim.x);
    const unsigned long long t0 = gtime();
    double acc = 0;
    for (size_t i = tid; i < nA; i += nt) { const double2 v = A[i]; acc += v.x + v.y; }
    if (acc == 1.2345) out[3] = 1;
    grid_bar(bar, 3 * gridDim.x);
    if (tid == 0) out[0] = gtime() - t0;
}
int main() {
    const size_t MB = 1 << 20;
    double2 *A, *B; unsigned *bar; unsigned long long *out;
    cudaMalloc(&A, 128 * MB); cudaMalloc(&B, 1024 * MB); cudaMalloc(&bar, 4); cudaMalloc(&out, 64);
    cudaMemset(B, 0, 1024 * MB);
    printf("%8s %10s | %10s %10s   (phase 3: read A back)\n", "A MB", "stream MB", "us", "GB/s");
    for (int a_mb : {12, 24, 48})
        for (int x_mb : {0, 8, 16, 32, 48, 64, 96, 128, 256}) {
            double best = 1e30;
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(bar, 0, 4);
                win<<<296, 256>>>(A, a_mb * MB / 16, B, (size_t)x_mb * MB / 32, bar, out, 0);
                unsigned long long t; cudaMemcpy(&t, out, 8, cudaMemcpyDeviceToHost);
                if (cudaGetLastError() != cudaSuccess) { printf("failed\n"); return 1; }
                best = t < best ? t : best;
            }
            printf("%8d %10d | %10.2f %10.0f\n", a_mb, x_mb, best / 1e3, a_mb * (double)MB / best);
            fflush(stdout);
        }
    return 0;
}
