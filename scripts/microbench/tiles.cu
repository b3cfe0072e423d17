// How the tile shape of a z-marching stencil kernel affects HBM bandwidth (no arithmetic, no halos):
// a block owns TW x TH double2 lanes of the xy plane and walks ZC planes; 8 arrays read, 3 written.
// Grid of the bench workload: pitch 192 doubles (96 double2), 182 rows, 366 planes (2 field sets).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tiles tiles.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

constexpr int PITCH2 = 96, ROWS = 182, PLANES = 366;
constexpr int NR = 8, NW = 3;

struct Ptrs { const double2 *in[NR]; double2 *out[NW]; };

template <int TW, int TH>
__global__ void __launch_bounds__(TW * TH) march(Ptrs p, int zc, int tiles_x, int tiles_y) {
    const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, kc = blockIdx.x / (tiles_x * tiles_y);
    const int i = tx * TW + threadIdx.x % TW, j = ty * TH + threadIdx.x / TW;
    if (i >= PITCH2 || j >= ROWS) return;
    const int kb = kc * zc, ke = min(kb + zc, PLANES);
    size_t x = ((size_t)kb * ROWS + j) * PITCH2 + i;
    for (int k = kb; k < ke; ++k, x += (size_t)ROWS * PITCH2) {
        double2 acc = make_double2(0, 0);
#pragma unroll
        for (int r = 0; r < NR; ++r) { const double2 v = p.in[r][x]; acc.x += v.x; acc.y += v.y; }
#pragma unroll
        for (int w = 0; w < NW; ++w) p.out[w][x] = make_double2(acc.x + w, acc.y);
    }
}

template <int TW, int TH>
static void run(const Ptrs &p, int zc, int smem = 0) {
    const int tiles_x = (PITCH2 + TW - 1) / TW, tiles_y = (ROWS + TH - 1) / TH, nzc = (PLANES + zc - 1) / zc;
    const int grid = tiles_x * tiles_y * nzc;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) march<TW, TH><<<grid, TW * TH, smem>>>(p, zc, tiles_x, tiles_y);
    cudaEventRecord(a);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) march<TW, TH><<<grid, TW * TH, smem>>>(p, zc, tiles_x, tiles_y);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)(NR + NW) * PITCH2 * ROWS * PLANES * sizeof(double2) * reps;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march<TW, TH>, TW * TH, smem);
    printf("tile %3d x %2d lanes (%4d B rows), %3d planes/block, grid %6d, %d blocks/SM : %7.1f GB/s  %s\n", TW, TH, TW * 16, zc, grid, nb,
           bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = (size_t)PITCH2 * ROWS * PLANES;
    Ptrs p;
    for (int r = 0; r < NR; ++r) { double2 *q; cudaMalloc(&q, n * sizeof(double2)); cudaMemset(q, 0, n * sizeof(double2)); p.in[r] = q; }
    for (int w = 0; w < NW; ++w) cudaMalloc(&p.out[w], n * sizeof(double2));
    for (int zc : {1, 4, 16, 64}) {
        run<16, 16>(p, zc);
        run<32, 8>(p, zc);
        run<96, 2>(p, zc);
        run<96, 4>(p, zc);
    }
    // occupancy sensitivity: dynamic shared memory caps the resident blocks per SM
    cudaFuncSetAttribute(march<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int zc : {8, 16})
        for (int kb : {28, 36, 50, 70, 100, 200}) run<16, 16>(p, zc, kb * 1024);
    return 0;
}
