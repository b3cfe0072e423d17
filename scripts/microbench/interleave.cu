// Many-stream kernels (the dispersive E-pass reads ~20 arrays and writes ~9) run below the few-stream
// HBM peak.  Does interleaving the arrays at row granularity ([plane][row][array][x] instead of
// [array][plane][row][x]) recover it?  Same tile walk as tiles.cu; G arrays per interleave group.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o interleave interleave.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int PITCH2 = 96, ROWS = 182, PLANES = 366, TW = 16, TH = 16, ZC = 16;

template <int NR, int NW, int G>
__global__ void __launch_bounds__(256) march(const double2 *__restrict__ in, double2 *__restrict__ out, int tiles_x, int tiles_y) {
    const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, kc = blockIdx.x / (tiles_x * tiles_y);
    const int i = tx * TW + threadIdx.x % TW, j = ty * TH + threadIdx.x / TW;
    if (i >= PITCH2 || j >= ROWS) return;
    const int kb = kc * ZC, ke = min(kb + ZC, PLANES);
    // array a = group * G + member: element (k, j, i) at ((group * PLANES + k) * ROWS + j) * G * PITCH2 + member * PITCH2 + i
    const size_t gstride = (size_t)PLANES * ROWS * G * PITCH2, pl = (size_t)ROWS * G * PITCH2;
    size_t x = ((size_t)kb * ROWS + j) * G * PITCH2 + i;
    for (int k = kb; k < ke; ++k, x += pl) {
        double2 acc = make_double2(0, 0);
#pragma unroll
        for (int r = 0; r < NR; ++r) { const double2 v = in[(r / G) * gstride + (r % G) * PITCH2 + x]; acc.x += v.x; acc.y += v.y; }
#pragma unroll
        for (int w = 0; w < NW; ++w) out[(w / G) * gstride + (w % G) * PITCH2 + x] = make_double2(acc.x + w, acc.y);
    }
}

template <int NR, int NW, int G>
static void run() {
    const size_t per = (size_t)PLANES * ROWS * PITCH2;
    double2 *in, *out;
    cudaMalloc(&in, ((NR + G - 1) / G) * G * per * sizeof(double2)); cudaMemset(in, 0, ((NR + G - 1) / G) * G * per * sizeof(double2));
    cudaMalloc(&out, ((NW + G - 1) / G) * G * per * sizeof(double2));
    const int tiles_x = (PITCH2 + TW - 1) / TW, tiles_y = (ROWS + TH - 1) / TH, nzc = (PLANES + ZC - 1) / ZC;
    const int grid = tiles_x * tiles_y * nzc;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) march<NR, NW, G><<<grid, 256>>>(in, out, tiles_x, tiles_y);
    cudaEventRecord(a);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) march<NR, NW, G><<<grid, 256>>>(in, out, tiles_x, tiles_y);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)(NR + NW) * per * sizeof(double2) * reps;
    printf("reads %2d writes %2d, %d arrays interleaved per row : %7.1f GB/s  %s\n", NR, NW, G, bytes / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(in); cudaFree(out);
}

int main() {
    run<8, 3, 1>(); run<8, 3, 3>();
    run<14, 6, 1>(); run<14, 6, 3>(); run<14, 6, 6>();
    run<20, 9, 1>(); run<20, 9, 3>(); run<20, 9, 6>();
    return 0;
}
