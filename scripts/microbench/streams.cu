// Practical HBM bandwidth of a streaming kernel as a function of the number of concurrent array streams.
// The FDTD passes read 8-26 arrays and write 3-10 per cell; this measures what a perfectly coalesced kernel
// with that many streams reaches on the same GPU (context for roofline.frac in DESIGN.md section 5).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o streams streams.cu ; run: ./streams
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int NR, int NW>
__global__ void __launch_bounds__(256) k(const double2 *const *__restrict__ in, double2 *const *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 acc = make_double2(0, 0);
#pragma unroll
        for (int r = 0; r < NR; ++r) { const double2 v = in[r][i]; acc.x += v.x; acc.y += v.y; }
#pragma unroll
        for (int w = 0; w < NW; ++w) out[w][i] = make_double2(acc.x + w, acc.y);
    }
}

template <int NR, int NW>
static void run(size_t n, int blocks_per_sm) {
    std::vector<double2 *> hin(NR), hout(NW);
    for (auto &p : hin) { cudaMalloc(&p, n * sizeof(double2)); cudaMemset(p, 0, n * sizeof(double2)); }
    for (auto &p : hout) cudaMalloc(&p, n * sizeof(double2));
    double2 **din, **dout;
    cudaMalloc(&din, NR * sizeof(void *)); cudaMalloc(&dout, NW * sizeof(void *));
    cudaMemcpy(din, hin.data(), NR * sizeof(void *), cudaMemcpyHostToDevice);
    cudaMemcpy(dout, hout.data(), NW * sizeof(void *), cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * blocks_per_sm;
    for (int w = 0; w < 3; ++w) k<NR, NW><<<grid, 256>>>(din, dout, n);
    cudaEventRecord(a);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) k<NR, NW><<<grid, 256>>>(din, dout, n);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)(NR + NW) * n * sizeof(double2) * reps;
    printf("reads %2d writes %2d  blocks/SM %d : %7.1f GB/s  (%s)\n", NR, NW, blocks_per_sm, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    for (auto p : hin) cudaFree(p);
    for (auto p : hout) cudaFree(p);
    cudaFree(din); cudaFree(dout);
}

int main() {
    const size_t n = (size_t)6 << 20;   // 96 MiB per array: far larger than L2
    run<1, 1>(n * 4, 8);
    run<2, 1>(n * 2, 8);
    run<8, 3>(n, 4); run<8, 3>(n, 8);
    run<11, 4>(n, 4); run<11, 4>(n, 8);
    run<23, 9>(n, 2); run<23, 9>(n, 4);
    return 0;
}
