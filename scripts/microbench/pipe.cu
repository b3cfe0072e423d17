// Does issuing the loads of several planes before the first use (software pipelining) replace occupancy?
// Same access pattern as tiles.cu (tile 16 x 16 double2 lanes, 8 arrays read, 3 written, 16 planes per block);
// UNR planes are loaded back to back before any of them is consumed; dynamic shared memory caps blocks/SM.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int PITCH2 = 96, ROWS = 182, PLANES = 366;
constexpr int NR = 8, NW = 3, TW = 16, TH = 16, ZC = 16;
struct Ptrs { const double2 *in[NR]; double2 *out[NW]; };

template <int UNR>
__global__ void __launch_bounds__(256) march(Ptrs p, int tiles_x, int tiles_y) {
    const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, kc = blockIdx.x / (tiles_x * tiles_y);
    const int i = tx * TW + threadIdx.x % TW, j = ty * TH + threadIdx.x / TW;
    if (i >= PITCH2 || j >= ROWS) return;
    const int kb = kc * ZC, ke = min(kb + ZC, PLANES);
    const size_t pl = (size_t)ROWS * PITCH2;
    size_t x = ((size_t)kb * ROWS + j) * PITCH2 + i;
    for (int k = kb; k < ke; k += UNR, x += UNR * pl) {
        double2 v[UNR][NR];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int r = 0; r < NR; ++r) if (k + u < ke) v[u][r] = p.in[r][x + u * pl];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (k + u >= ke) break;
            double2 acc = make_double2(0, 0);
#pragma unroll
            for (int r = 0; r < NR; ++r) { acc.x += v[u][r].x; acc.y += v[u][r].y; }
#pragma unroll
            for (int w = 0; w < NW; ++w) p.out[w][x + u * pl] = make_double2(acc.x + w, acc.y);
        }
    }
}

template <int UNR>
static void run(const Ptrs &p, int smem_kb) {
    const int tiles_x = (PITCH2 + TW - 1) / TW, tiles_y = (ROWS + TH - 1) / TH, nzc = (PLANES + ZC - 1) / ZC;
    const int grid = tiles_x * tiles_y * nzc, smem = smem_kb * 1024;
    cudaFuncSetAttribute(march<UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, march<UNR>);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) march<UNR><<<grid, 256, smem>>>(p, tiles_x, tiles_y);
    cudaEventRecord(a);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) march<UNR><<<grid, 256, smem>>>(p, tiles_x, tiles_y);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)(NR + NW) * PITCH2 * ROWS * PLANES * sizeof(double2) * reps;
    printf("planes in flight %d, regs %3d, smem cap %3d KB (~%d blocks/SM) : %7.1f GB/s  %s\n", UNR, fa.numRegs, smem_kb,
           smem_kb ? 227 / smem_kb : 8, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = (size_t)PITCH2 * ROWS * PLANES;
    Ptrs p;
    for (int r = 0; r < NR; ++r) { double2 *q; cudaMalloc(&q, n * sizeof(double2)); cudaMemset(q, 0, n * sizeof(double2)); p.in[r] = q; }
    for (int w = 0; w < NW; ++w) cudaMalloc(&p.out[w], n * sizeof(double2));
    for (int kb : {0, 50, 70, 100, 200}) { run<1>(p, kb); run<2>(p, kb); run<4>(p, kb); }
    return 0;
}
