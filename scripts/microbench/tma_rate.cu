// tma_rate.cu -- how fast can producer threads of one block per SM issue small TMA boxes?  (B200, sm_100a)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu && ./tma_rate
// One block per SM, W producer warps (one elected thread each), each with its own ring of S slots.  A "load" is K boxes of
// (bw x bh x bd) doubles armed on one mbarrier; the producer waits for the load issued S loads earlier before it reuses the slot
// (no consumers: this measures the TMA issue path and the memory system only).  Prints cycles per load / per op and GB/s.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// mapmode 0: descriptor in global memory (const CUtensorMap *), 1: __grid_constant__ kernel parameter
template <int MAPMODE>
__global__ void __launch_bounds__(256) rate(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int W, int S, int K, int slot_bytes, int box_bytes, int box_slot,
                                            int n_loads, int tiles_x, int tiles_y, int bw, int bh, int bd, int planes, int dram, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = smem_u32(smem);                // [W][S] barriers
    const uint32_t data = (bars + 8u * W * S + 127u) & ~127u;
    if (threadIdx.x == 0) { for (int i = 0; i < W * S; ++i) mbar_init(bars + 8u * i, 1u); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    __syncthreads();
    if (warp >= W || lane != 0) return;
    const CUtensorMap *m = MAPMODE ? &pmap : gmap;
    const int ring = warp * S;
    // every producer walks its own tile column through the planes
    // dram != 0: every producer walks a (column, z range) of its own, nothing is read twice (no L2 hits)
    const int g = blockIdx.x * W + warp, ncol = tiles_x * tiles_y;
    const int col = g % ncol;
    const int i0 = (col % tiles_x) * bw, j0 = (col / tiles_x) * bh;
    const int zspan = n_loads * K * bd;
    int z = dram ? (g / ncol) * zspan : (g * 37) % (planes - K * bd);
    const long long t0 = clock64();
    for (int n = 0; n < n_loads; ++n) {
        const int sl = n % S;
        const uint32_t bar = bars + 8u * (ring + sl);
        if (n >= S) mbar_wait(bar, (unsigned)((n / S - 1) & 1));
        mbar_expect_tx(bar, (uint32_t)(K * box_bytes));
        const uint32_t d = data + (uint32_t)((ring + sl) * slot_bytes);
        for (int c = 0; c < K; ++c) tma_load_3d(d + c * box_slot, m, i0, j0, z + c * bd, bar);
        z += K * bd; if (!dram && z > planes - K * bd) z = 0;
    }
    const long long t1 = clock64();
    for (int n = (n_loads > S ? n_loads - S : 0); n < n_loads; ++n) mbar_wait(bars + 8u * (ring + n % S), (unsigned)((n / S) & 1));
    const long long t2 = clock64();
    out[(blockIdx.x * 8 + warp) * 2] = t1 - t0; out[(blockIdx.x * 8 + warp) * 2 + 1] = t2 - t0;
}

int main(int argc, char **argv) {
    void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int pitch = 192, rows = 182, planes = 34000;     // 9.5 GB: room for distinct regions
    double *buf; cudaMalloc(&buf, (size_t)pitch * rows * planes * 8); cudaMemset(buf, 0, (size_t)pitch * rows * planes * 8);
    long long *out; cudaMalloc(&out, 148 * 8 * 2 * 8);
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-30s %4s %5s %3s %3s %3s | %9s %9s %9s %9s\n", "box (doubles)", "src", "map", "W", "S", "K", "cyc/load", "cyc/op", "GB/s", "B/clk/SM");
    struct Case { int bw, bh, bd, K, W, S, mode, dram; };
    std::vector<Case> cases;
    for (int dram = 0; dram < 2; ++dram) {
        cases.push_back({54, 9, 1, 6, 1, 6, 0, dram});       // the H-pass interior load (6 boxes of 3.9 KB), 6 loads in flight
        cases.push_back({54, 17, 1, 6, 1, 3, 0, dram});      // tall tiles: 6 boxes of 7.3 KB, 3 loads in flight
        cases.push_back({54, 17, 1, 6, 1, 4, 0, dram});
        cases.push_back({54, 33, 1, 6, 1, 2, 0, dram});
        cases.push_back({54, 9, 1, 6, 1, 8, 0, dram});
        cases.push_back({54, 9, 1, 6, 2, 4, 0, dram});       // two producer warps
        cases.push_back({54, 17, 1, 6, 2, 2, 0, dram});
        cases.push_back({16, 1, 1, 6, 1, 16, 0, dram});      // tiny boxes: the per-operation cost
    }
    for (const Case &c : cases) {
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * rows * 8};
        const cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, (cuuint32_t)c.bd}, es[3] = {1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
        cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice);
        const int box_bytes = c.bw * c.bh * c.bd * 8, box_slot = (box_bytes + 127) / 128 * 128, slot = c.K * box_slot;
        const size_t smem = 8 * c.W * c.S + 128 + (size_t)c.W * c.S * slot;
        if (smem > 227 * 1024) { printf("skip (smem %zu)\n", smem); continue; }
        const int n_loads = 400;
        auto kern = c.mode ? rate<1> : rate<0>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int tw = c.bw - 2, th = c.bh > 1 ? c.bh - 1 : 1;
        for (int rep = 0; rep < 2; ++rep)
            kern<<<148, 256, smem>>>(map, gmap, c.W, c.S, c.K, slot, box_bytes, box_slot, n_loads, 181 / tw, 181 / th, tw, th, c.bd, planes, c.dram, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<long long> h(148 * 8 * 2);
        cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
        double issue = 0, total = 0;
        for (int b = 0; b < 148; ++b) for (int w = 0; w < c.W; ++w) { issue += h[(b * 8 + w) * 2]; total += h[(b * 8 + w) * 2 + 1]; }
        issue /= 148.0 * c.W; total /= 148.0 * c.W;
        const double bytes_sm = (double)n_loads * c.K * box_bytes * c.W;
        char name[64]; snprintf(name, sizeof name, "%d x %d x %d (%d B)", c.bw, c.bh, c.bd, box_bytes);
        printf("%-30s %4s %5s %3d %3d %3d | %9.0f %9.0f %9.0f %9.1f\n", name, c.dram ? "dram" : "mix", c.mode ? "param" : "gmem", c.W, c.S, c.K, total / n_loads, total / n_loads / c.K,
               bytes_sm * 148 / (total / (clk * 1e3)) / 1e9, bytes_sm / total);
        fflush(stdout);
    }
    return 0;
}
