// sj_tma.cuh -- TMA-staged column kernels of the FDTD hot path (sm_100a).
//
// One persistent kernel per (half-pass, stage layout).  A thread block owns a list of work items (xy tile x run of
// z planes, any cell class: interior, single-sigma PML face, PML edge / corner) fixed by the host (LPT schedule), and
// marches each item along z.  Data movement is Blackwell's: one elected producer thread issues a
// cp.async.bulk.tensor (TMA) box load per array per plane into a ring of shared-memory stages and signals an mbarrier
// with the byte count; the consumer warps wait on that barrier, read their cells and the i/j neighbours from the staged
// boxes (the curl inputs are loaded with a one-vector / one-row halo, so there are no shuffles, no tile-edge loads and
// no boundary predicates: TMA zero-fills outside the grid), do the arithmetic of sj_kernels.cuh and store with 128-bit
// st.global.  The ring is released stage by stage through a second set of mbarriers (one arrival per consumer warp), so
// the producer runs NST-1 planes ahead -- also across the boundary between two items -- whatever the occupancy.
//
// The arithmetic (curl order, UPML forms, ADE) is expression for expression that of the register kernels in
// sj_kernels.cuh, so both paths give bit-identical fields (tests/test_gpu_tma.py).
#pragma once
#include <cuda.h>

#include "sj_kernels.cuh"

// tensor maps per tile shape: F with halo, F own cells, P own cells, then one per PML box (own cells of its 12 arrays)
#define SJ_TMAP_F_HALO 0
#define SJ_TMAP_F_OWN 1
#define SJ_TMAP_P_OWN 2
#define SJ_TMAP_BOX0 3
#define SJ_TMAP_PER_SHAPE (3 + SJ_N_PML_BOX)

struct TmaPlan {
    const CUtensorMap *maps;              // [shape][SJ_TMAP_PER_SHAPE], device memory
    TShape shape[SJ_TMA_MAX_SHAPES];
    const WorkItem *items;
    const int *blk_first;                 // [gridDim.x + 1]: block b owns items [blk_first[b], blk_first[b+1])
};

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <typename T, int V>
__device__ __forceinline__ void lds_vec(Vec<T, V> &v, const T *p) { v.load(p); }

// ring bookkeeping shared by producer and consumers: load number q -> stage q % NST, phase (q / NST) & 1
template <int NST>
struct Ring {
    uint32_t full0, empty0, data0;      // shared-space addresses (barriers, stage 0)
    unsigned char *data_gen;            // generic address of stage 0
    int stage_bytes;
    __device__ __forceinline__ uint32_t full(int q) const { return full0 + 8u * (unsigned)(q % NST); }
    __device__ __forceinline__ uint32_t empty(int q) const { return empty0 + 8u * (unsigned)(q % NST); }
    __device__ __forceinline__ uint32_t phase(int q) const { return (unsigned)(q / NST) & 1u; }
    __device__ __forceinline__ uint32_t data(int q) const { return data0 + (unsigned)(q % NST) * (unsigned)stage_bytes; }
    __device__ __forceinline__ const unsigned char *gen(int q) const { return data_gen + (q % NST) * stage_bytes; }
};

// Stage layout (byte offsets inside one stage), NT = consumer threads:
//   3 curl-input boxes with halo (HALO bytes each), 3 own-field tiles, NAUX auxiliary tiles, NPOL polarisation tiles
template <int NT> struct Slots { static constexpr int OWN = NT * 16, HALO = NT * 20; };

template <int NST>
__device__ __forceinline__ void ring_setup(Ring<NST> &r, unsigned char *smem, int stage_bytes, int n_consumer_warps) {
    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    r.full0 = base; r.empty0 = base + 8u * NST; r.data0 = base + 128u; r.stage_bytes = stage_bytes;
    r.data_gen = smem + (base + 128u - smem_u32(smem));
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(r.full0 + 8u * s, 1u); mbar_init(r.empty0 + 8u * s, (unsigned)n_consumer_warps); }
        mbar_fence_init();
    }
    __syncthreads();
}

// z coordinate (dimension 2 of the tensor maps) of plane kl of (array a, set) for arrays stored [a][set][plane]
__device__ __forceinline__ int zcoord(int a, int n_sets, int set, int planes, int kl) { return (a * n_sets + set) * planes + kl; }

// =====================================================================================================
// H-pass
// =====================================================================================================
// aux slots: class A (GENERAL = false): 1 (the normal B of a face tile; unused by interior tiles);
//            GENERAL: 6 (Bx, By, Bz, Ux, Uy, Uz)
template <typename T, int NT, int NST, bool GENERAL>
__device__ __forceinline__ void h_tma_produce(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const Ring<NST> &r,
                                              int k_lo, int k_hi) {
    typedef Slots<NT> SL;
    int q = 0;
    const int n0 = plan.blk_first[blockIdx.x], n1 = plan.blk_first[blockIdx.x + 1];
    for (int n = n0; n < n1; ++n) {
        const WorkItem it = plan.items[n];
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        if (kb >= ke) continue;
        const TShape sh = plan.shape[it.shape];
        const CUtensorMap *mp = plan.maps + it.shape * SJ_TMAP_PER_SHAPE;
        const uint32_t hb = (uint32_t)(sh.hp * (sh.th + 1)) * sizeof(T), ob = (uint32_t)(sh.tw * sh.th) * sizeof(T);
        const int naux = GENERAL ? 6 : (it.kind != 0 && it.box >= 0 ? 1 : 0);
        int bi = 0, bj = 0, bz = 1, bk0 = 0;
        const CUtensorMap *mb = mp;
        if (it.box >= 0) {
            const PmlBox<T> &b = bs.b[it.box];
            bi = it.i0 - b.lo[0]; bj = it.j0 - b.lo[1]; bz = b.hi[2] - b.lo[2]; bk0 = b.lo[2];
            mb = mp + SJ_TMAP_BOX0 + it.box;
        }
        for (int k = kb; k <= ke; ++k, ++q) {
            mbar_wait(r.empty(q), r.phase(q) ^ 1u);
            const uint32_t bar = r.full(q), d = r.data(q);
            const int kl = k - p.kz0 + 1;
            if (k == ke) {               // the plane above the run: Ex, Ey only
                mbar_expect_tx(bar, 2 * hb);
                tma_load_3d(d, mp + SJ_TMAP_F_HALO, it.i0, it.j0, zcoord(0, p.n_sets, it.set, p.nzl, kl), bar);
                tma_load_3d(d + SL::HALO, mp + SJ_TMAP_F_HALO, it.i0, it.j0, zcoord(1, p.n_sets, it.set, p.nzl, kl), bar);
                continue;
            }
            mbar_expect_tx(bar, 3 * hb + (3 + naux) * ob);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                tma_load_3d(d + c * SL::HALO, mp + SJ_TMAP_F_HALO, it.i0, it.j0, zcoord(c, p.n_sets, it.set, p.nzl, kl), bar);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                tma_load_3d(d + 3 * SL::HALO + c * SL::OWN, mp + SJ_TMAP_F_OWN, it.i0, it.j0, zcoord(3 + c, p.n_sets, it.set, p.nzl, kl), bar);
            if (GENERAL) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {     // B = array group 1, UB = array group 3 of the box allocation
                    tma_load_3d(d + 3 * SL::HALO + (3 + c) * SL::OWN, mb, bi, bj, zcoord(3 + c, p.n_sets, it.set, bz, k - bk0), bar);
                    tma_load_3d(d + 3 * SL::HALO + (6 + c) * SL::OWN, mb, bi, bj, zcoord(9 + c, p.n_sets, it.set, bz, k - bk0), bar);
                }
            } else if (naux) {
                tma_load_3d(d + 3 * SL::HALO + 3 * SL::OWN, mb, bi, bj, zcoord(3 + (it.kind - 1), p.n_sets, it.set, bz, k - bk0), bar);
            }
        }
    }
}

// PD: 4 interior, 1/2/3 face with normal x/y/z, 0 general
template <typename T, int NT, int NST, int PD>
__device__ __forceinline__ void h_tma_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const WorkItem &it, const TShape &sh,
                                           const Ring<NST> &r, int kb, int ke, int &q) {
    typedef Slots<NT> SL;
    constexpr int V = 16 / (int)sizeof(T);
    const int t = threadIdx.x, lane = t & 31;
    const int row = t / sh.nvx, vx = t - row * sh.nvx;
    const bool in_tile = row < sh.th;
    const int i0 = it.i0 + vx * V, j = it.j0 + row;
    const bool act = in_tile && (j < it.j_hi) && (i0 < it.i_hi);
    const int hc = in_tile ? row * sh.hp + vx * V : 0;          // element offsets inside a halo box / an own tile
    const int hj = in_tile ? hc + sh.hp : 0;
    const int oc = in_tile ? row * sh.tw + vx * V : 0;
    const T C = p.courant;
    const long long plane = p.plane;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    T *pH = p.F + 3 * fcs + (long long)it.set * p.set_stride + (long long)(kb - p.kz0 + 1) * plane + (long long)j * p.pitch + i0;
    long long bcs = 0, bplane = 0;
    T *pB = nullptr, *pU = nullptr;
    if (PD != 4) {
        const PmlBox<T> &b = bs.b[it.box];
        bcs = b.bcs; bplane = b.bplane;
        const long long xb0 = (long long)it.set * b.bset + (long long)(kb - b.lo[2]) * bplane + (long long)(j - b.lo[1]) * b.bpitch + (i0 - b.lo[0]);
        pB = b.B[0] + xb0; pU = b.UB[0] + xb0;
    }
    T sxi[V], sxh[V], ixh[V];
    T syi = T(0), syh = T(0), iyh = T(1);
#pragma unroll
    for (int v = 0; v < V; ++v) { sxi[v] = T(0); sxh[v] = T(0); ixh[v] = T(1); }
    if (act) {
        if (PD == 0 || PD == 1) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int h = 2 * (i0 + v);
                sxi[v] = p.sig[0][h]; sxh[v] = p.sig[0][h + 1]; ixh[v] = p.siginv[0][h + 1];
            }
        }
        if (PD == 0 || PD == 2) { syi = p.sig[1][2 * j]; syh = p.sig[1][2 * j + 1]; iyh = p.siginv[1][2 * j + 1]; }
    }
    const bool jok = (j <= p.n[1] - 1);

    Vec<T, V> ex0, ey0, ex1, ey1, ez0, ezj, exj, hx, hy, hz, bx, by, bz, ux, uy, uz;
    bx.zero(); by.zero(); bz.zero(); ux.zero(); uy.zero(); uz.zero();
    mbar_wait(r.full(q), r.phase(q));
    {
        const unsigned char *d = r.gen(q);
        ex0.load(reinterpret_cast<const T *>(d) + hc); ey0.load(reinterpret_cast<const T *>(d + SL::HALO) + hc);
    }
    for (int k = kb; k < ke; ++k, ++q) {
        mbar_wait(r.full(q + 1), r.phase(q + 1));
        const unsigned char *d0 = r.gen(q), *d1 = r.gen(q + 1);
        const T *sEx = reinterpret_cast<const T *>(d0), *sEy = reinterpret_cast<const T *>(d0 + SL::HALO), *sEz = reinterpret_cast<const T *>(d0 + 2 * SL::HALO);
        const T *sH = reinterpret_cast<const T *>(d0 + 3 * SL::HALO);
        constexpr int OWN_E = SL::OWN / (int)sizeof(T);
        ex1.load(reinterpret_cast<const T *>(d1) + hc); ey1.load(reinterpret_cast<const T *>(d1 + SL::HALO) + hc);
        ez0.load(sEz + hc); ezj.load(sEz + hj); exj.load(sEx + hj);
        const T ez_n = sEz[hc + V], ey_n = sEy[hc + V];
        hx.load(sH + oc); hy.load(sH + OWN_E + oc); hz.load(sH + 2 * OWN_E + oc);
        if (PD == 0) {
            bx.load(sH + 3 * OWN_E + oc); by.load(sH + 4 * OWN_E + oc); bz.load(sH + 5 * OWN_E + oc);
            ux.load(sH + 6 * OWN_E + oc); uy.load(sH + 7 * OWN_E + oc); uz.load(sH + 8 * OWN_E + oc);
        } else if (PD == 1) bx.load(sH + 3 * OWN_E + oc);
        else if (PD == 2) by.load(sH + 3 * OWN_E + oc);
        else if (PD == 3) bz.load(sH + 3 * OWN_E + oc);
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty(q));           // stage of plane k is free (plane k+1 stays until the next iteration)
        if (act) {
            if (PD == 4) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                    const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                    hx.v[v] -= C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1.v[v]);
                    hy.v[v] -= C * (((ex1.v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
                    hz.v[v] -= C * (((eyi - ey0.v[v]) + ex0.v[v]) - exj.v[v]);
                }
            } else {
                T szi = T(0), szh = T(0), izh = T(1);
                if (PD == 0 || PD == 3) { szi = p.sig[2][2 * k]; szh = p.sig[2][2 * k + 1]; izh = p.siginv[2][2 * k + 1]; }
                const bool kok = (k <= p.n[2] - 1);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int i = i0 + v;
                    const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                    const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                    const bool iok = (i <= p.n[0] - 1);
                    const T sf = PD == 1 ? sxh[v] : PD == 2 ? syh : szh, isf = PD == 1 ? ixh[v] : PD == 2 ? iyh : izh;
                    if (i >= 1 && iok && jok && kok) {       // Hx: k-dir y, u-dir z, w-dir x
                        const T curl = C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1.v[v]);
                        if (PD == 0) {
                            const T bo = bx.v[v];
                            const T bn = pml_step_db(bo, curl, syh, iyh, szh, izh, ux.v[v]);
                            bx.v[v] = bn;
                            hx.v[v] = (sxi[v] != T(0)) ? hx.v[v] + (T(1) + sxi[v]) * bn - (T(1) - sxi[v]) * bo : bn;
                        } else if (PD == 1) {
                            const T bo = bx.v[v], bn = bo - curl;
                            bx.v[v] = bn;
                            hx.v[v] += (T(1) + sxi[v]) * bn - (T(1) - sxi[v]) * bo;
                        } else hx.v[v] = ((T(1) - sf) * hx.v[v] - curl) * isf;
                    }
                    if (j >= 1 && jok && iok && kok) {       // Hy: k-dir z, u-dir x, w-dir y
                        const T curl = C * (((ex1.v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
                        if (PD == 0) {
                            const T bo = by.v[v];
                            const T bn = pml_step_db(bo, curl, szh, izh, sxh[v], ixh[v], uy.v[v]);
                            by.v[v] = bn;
                            hy.v[v] = (syi != T(0)) ? hy.v[v] + (T(1) + syi) * bn - (T(1) - syi) * bo : bn;
                        } else if (PD == 2) {
                            const T bo = by.v[v], bn = bo - curl;
                            by.v[v] = bn;
                            hy.v[v] += (T(1) + syi) * bn - (T(1) - syi) * bo;
                        } else hy.v[v] = ((T(1) - sf) * hy.v[v] - curl) * isf;
                    }
                    if (k >= 1 && kok && iok && jok) {       // Hz: k-dir x, u-dir y, w-dir z
                        const T curl = C * (((eyi - ey0.v[v]) + ex0.v[v]) - exj.v[v]);
                        if (PD == 0) {
                            const T bo = bz.v[v];
                            const T bn = pml_step_db(bo, curl, sxh[v], ixh[v], syh, iyh, uz.v[v]);
                            bz.v[v] = bn;
                            hz.v[v] = (szi != T(0)) ? hz.v[v] + (T(1) + szi) * bn - (T(1) - szi) * bo : bn;
                        } else if (PD == 3) {
                            const T bo = bz.v[v], bn = bo - curl;
                            bz.v[v] = bn;
                            hz.v[v] += (T(1) + szi) * bn - (T(1) - szi) * bo;
                        } else hz.v[v] = ((T(1) - sf) * hz.v[v] - curl) * isf;
                    }
                }
            }
            hx.store(pH); hy.store(pH + fcs); hz.store(pH + fcs2);
            if (PD == 0 || PD == 1) bx.store(pB);
            if (PD == 0 || PD == 2) by.store(pB + bcs);
            if (PD == 0 || PD == 3) bz.store(pB + 2 * bcs);
            if (PD == 0) { ux.store(pU); uy.store(pU + bcs); uz.store(pU + 2 * bcs); }
        }
        ex0 = ex1; ey0 = ey1;
        pH += plane; pB += bplane; pU += bplane;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(r.empty(q));               // the Ex, Ey plane above the run
    ++q;
}

template <typename T, int NT, int NST, bool GENERAL, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) h_tma(const KParams<T> p, const PmlBoxSet<T> bs, const TmaPlan plan,
                                                                  int k_lo, int k_hi) {
    typedef Slots<NT> SL;
    extern __shared__ unsigned char sj_tma_smem[];
    Ring<NST> r;
    ring_setup<NST>(r, sj_tma_smem, 3 * SL::HALO + (3 + (GENERAL ? 6 : 1)) * SL::OWN, NT / 32);
    if (threadIdx.x >= NT) {
        if (threadIdx.x == NT) h_tma_produce<T, NT, NST, GENERAL>(p, bs, plan, r, k_lo, k_hi);
        return;
    }
    int q = 0;
    const int n0 = plan.blk_first[blockIdx.x], n1 = plan.blk_first[blockIdx.x + 1];
    for (int n = n0; n < n1; ++n) {
        const WorkItem it = plan.items[n];
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        if (kb >= ke) continue;
        const TShape sh = plan.shape[it.shape];
        if (GENERAL) h_tma_item<T, NT, NST, 0>(p, bs, it, sh, r, kb, ke, q);
        else if (it.box < 0) h_tma_item<T, NT, NST, 4>(p, bs, it, sh, r, kb, ke, q);
        else if (it.kind == 1) h_tma_item<T, NT, NST, 1>(p, bs, it, sh, r, kb, ke, q);
        else if (it.kind == 2) h_tma_item<T, NT, NST, 2>(p, bs, it, sh, r, kb, ke, q);
        else h_tma_item<T, NT, NST, 3>(p, bs, it, sh, r, kb, ke, q);
    }
}

// =====================================================================================================
// E-pass
// =====================================================================================================
// Stage: Hx, Hy, Hz boxes with the low-side halo (origin i0 - V, j0 - 1) | Ex, Ey, Ez tiles | NAUX auxiliary tiles
// (class A: the normal D of a face tile; GENERAL: Dx, Dy, Dz, Ux, Uy, Uz) | polarisation tiles [c][s]{current, previous}.
// NS = pole slots staged (0: none); UNI: the item holds one material (it.mat) -- chi, eps and the folded ADE coefficients
// sit in registers; otherwise (mixed tiles) the material bytes are read with plain loads one plane ahead.
template <int NT, int NS, bool GENERAL> struct EStage {
    typedef Slots<NT> SL;
    static constexpr int NAUX = GENERAL ? 6 : 1;
    static constexpr int OFF_E = 3 * SL::HALO, OFF_AUX = OFF_E + 3 * SL::OWN, OFF_P = OFF_AUX + NAUX * SL::OWN;
    static constexpr int BYTES = OFF_P + 6 * NS * SL::OWN;
};

template <typename T, int NT, int NST, int NS, bool GENERAL>
__device__ __forceinline__ void e_tma_produce(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const Ring<NST> &r,
                                              int k_lo, int k_hi) {
    typedef Slots<NT> SL;
    typedef EStage<NT, NS, GENERAL> ES;
    constexpr int V = 16 / (int)sizeof(T);
    int q = 0;
    const int parity = (int)(*p.step & 1);
    const int n0 = plan.blk_first[blockIdx.x], n1 = plan.blk_first[blockIdx.x + 1];
    for (int n = n0; n < n1; ++n) {
        const WorkItem it = plan.items[n];
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        if (kb >= ke) continue;
        const TShape sh = plan.shape[it.shape];
        const CUtensorMap *mp = plan.maps + it.shape * SJ_TMAP_PER_SHAPE;
        const uint32_t hb = (uint32_t)(sh.hp * (sh.th + 1)) * sizeof(T), ob = (uint32_t)(sh.tw * sh.th) * sizeof(T);
        const int naux = GENERAL ? 6 : (it.kind != 0 && it.box >= 0 ? 1 : 0);
        int bi = 0, bj = 0, bz = 1, bk0 = 0;
        const CUtensorMap *mb = mp;
        if (it.box >= 0) {
            const PmlBox<T> &b = bs.b[it.box];
            bi = it.i0 - b.lo[0]; bj = it.j0 - b.lo[1]; bz = b.hi[2] - b.lo[2]; bk0 = b.lo[2];
            mb = mp + SJ_TMAP_BOX0 + it.box;
        }
        const int hi0 = it.i0 - V, hj0 = it.j0 - 1;
        for (int k = kb - 1; k < ke; ++k, ++q) {
            mbar_wait(r.empty(q), r.phase(q) ^ 1u);
            const uint32_t bar = r.full(q), d = r.data(q);
            const int kl = k - p.kz0 + 1;
            if (k == kb - 1) {           // the plane below the run: Hx, Hy only
                mbar_expect_tx(bar, 2 * hb);
                tma_load_3d(d, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(3, p.n_sets, it.set, p.nzl, kl), bar);
                tma_load_3d(d + SL::HALO, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(4, p.n_sets, it.set, p.nzl, kl), bar);
                continue;
            }
            mbar_expect_tx(bar, 3 * hb + (3 + naux + 6 * NS) * ob);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                tma_load_3d(d + c * SL::HALO, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(3 + c, p.n_sets, it.set, p.nzl, kl), bar);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                tma_load_3d(d + ES::OFF_E + c * SL::OWN, mp + SJ_TMAP_F_OWN, it.i0, it.j0, zcoord(c, p.n_sets, it.set, p.nzl, kl), bar);
            if (GENERAL) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {     // D = array group 0, UD = array group 2 of the box allocation
                    tma_load_3d(d + ES::OFF_AUX + c * SL::OWN, mb, bi, bj, zcoord(c, p.n_sets, it.set, bz, k - bk0), bar);
                    tma_load_3d(d + ES::OFF_AUX + (3 + c) * SL::OWN, mb, bi, bj, zcoord(6 + c, p.n_sets, it.set, bz, k - bk0), bar);
                }
            } else if (naux) {
                tma_load_3d(d + ES::OFF_AUX, mb, bi, bj, zcoord(it.kind - 1, p.n_sets, it.set, bz, k - bk0), bar);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int ac = (parity * p.n_slots + s) * 3 + c, ap = ((parity ^ 1) * p.n_slots + s) * 3 + c;
                    tma_load_3d(d + ES::OFF_P + ((c * NS + s) * 2) * SL::OWN, mp + SJ_TMAP_P_OWN, it.i0, it.j0, zcoord(ac, p.n_sets, it.set, p.nzl, kl), bar);
                    tma_load_3d(d + ES::OFF_P + ((c * NS + s) * 2 + 1) * SL::OWN, mp + SJ_TMAP_P_OWN, it.i0, it.j0, zcoord(ap, p.n_sets, it.set, p.nzl, kl), bar);
                }
        }
    }
}

template <typename T, int NT, int NST, int NS, int PD, bool SRC, bool UNI, bool GENERAL>
__device__ __forceinline__ void e_tma_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const WorkItem &it, const TShape &sh,
                                           const Ring<NST> &r, int kb, int ke, int &q) {
    typedef Slots<NT> SL;
    typedef EStage<NT, NS, GENERAL> ES;
    constexpr int V = 16 / (int)sizeof(T);
    constexpr bool POL = NS > 0;
    constexpr bool GEN = POL && !UNI;             // per-cell material bytes and table look-ups
    constexpr int NS1 = NS > 0 ? NS : 1;
    constexpr int OWN_E = SL::OWN / (int)sizeof(T);
    const int t = threadIdx.x, lane = t & 31;
    const int row = t / sh.nvx, vx = t - row * sh.nvx;
    const bool in_tile = row < sh.th;
    const int i0 = it.i0 + vx * V, j = it.j0 + row;
    const bool act = in_tile && (j < it.j_hi) && (i0 < it.i_hi);
    const int hcen = in_tile ? (row + 1) * sh.hp + (vx + 1) * V : V;    // own cell inside the halo box (origin i0 - V, j0 - 1)
    const int hjm = in_tile ? hcen - sh.hp : V;
    const int oc = in_tile ? row * sh.tw + vx * V : 0;
    const T C = p.courant;
    const int set = it.set;
    const long long step = *p.step;
    const int parity = (int)(step & 1);
    const long long plane = p.plane;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    const long long xl0 = (long long)(kb - p.kz0 + 1) * plane + (long long)j * p.pitch + i0;
    long long xg = (long long)set * p.set_stride + xl0;
    T *pE = p.F + xg;
    long long bcs = 0, bplane = 0;
    T *pD = nullptr, *pU = nullptr;
    if (PD != 4) {
        const PmlBox<T> &b = bs.b[it.box];
        bcs = b.bcs; bplane = b.bplane;
        const long long xb0 = (long long)set * b.bset + (long long)(kb - b.lo[2]) * bplane + (long long)(j - b.lo[1]) * b.bpitch + (i0 - b.lo[0]);
        pD = b.D[0] + xb0; pU = b.UD[0] + xb0;
    }
    const long long mcs = p.set_stride, mcs2 = 2 * p.set_stride;
    const uint8_t *pm = p.mat[0] + xl0;
    const long long pcs = p.p_comp_stride;
    T *bprv = p.Pall + (long long)(parity ^ 1) * p.n_slots * 3 * pcs;          // read as "previous", written as new

    T sxi[V], ixi[V], sxh[V];
    T syi = T(0), iyi = T(1), syh = T(0);
#pragma unroll
    for (int v = 0; v < V; ++v) { sxi[v] = T(0); ixi[v] = T(1); sxh[v] = T(0); }
    if (act) {
        if (PD == 0 || PD == 1) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int h = 2 * (i0 + v);
                sxi[v] = p.sig[0][h]; ixi[v] = p.siginv[0][h]; sxh[v] = p.sig[0][h + 1];
            }
        }
        if (PD == 0 || PD == 2) { syi = p.sig[1][2 * j]; iyi = p.siginv[1][2 * j]; syh = p.sig[1][2 * j + 1]; }
    }
    const bool jin = (j >= 1 && j <= p.n[1] - 1);
    T chi_u = T(0), eps_u = T(0), cfu[NS1][3];
#pragma unroll
    for (int s_ = 0; s_ < NS1; ++s_)
#pragma unroll
        for (int q_ = 0; q_ < 3; ++q_) cfu[s_][q_] = (UNI && POL) ? p.mt_coef[((long long)it.mat * SJ_MAX_POLES + s_) * 3 + q_] : T(0);
    if (UNI) { chi_u = p.mt_chi[it.mat]; eps_u = p.mt_eps[it.mat]; }

    Vec<T, V> hxm, hym, hx0, hy0, hz0, hzj, hxj, ex, ey, ez, dx, dy, dz, ux, uy, uz;
    dx.zero(); dy.zero(); dz.zero(); ux.zero(); uy.zero(); uz.zero();
    unsigned char mx[V], my[V], mz[V], nx_[V], ny_[V], nz_[V];
#pragma unroll
    for (int v = 0; v < V; ++v) mx[v] = my[v] = mz[v] = nx_[v] = ny_[v] = nz_[v] = 0;
    if (GEN && act) { load_bytes<V>(pm, mx); load_bytes<V>(pm + mcs, my); load_bytes<V>(pm + mcs2, mz); }
    PolState<T, V, NS> pol;

    mbar_wait(r.full(q), r.phase(q));
    {
        const unsigned char *d = r.gen(q);
        hxm.load(reinterpret_cast<const T *>(d) + hcen); hym.load(reinterpret_cast<const T *>(d + SL::HALO) + hcen);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(r.empty(q));
    ++q;
    for (int k = kb; k < ke; ++k, ++q) {
        if (GEN && act) { load_bytes<V>(pm + plane, nx_); load_bytes<V>(pm + mcs + plane, ny_); load_bytes<V>(pm + mcs2 + plane, nz_); }
        mbar_wait(r.full(q), r.phase(q));
        const unsigned char *d0 = r.gen(q);
        const T *sHx = reinterpret_cast<const T *>(d0), *sHy = reinterpret_cast<const T *>(d0 + SL::HALO), *sHz = reinterpret_cast<const T *>(d0 + 2 * SL::HALO);
        const T *sE = reinterpret_cast<const T *>(d0 + ES::OFF_E), *sA = reinterpret_cast<const T *>(d0 + ES::OFF_AUX);
        const T *sP = reinterpret_cast<const T *>(d0 + ES::OFF_P);
        hx0.load(sHx + hcen); hy0.load(sHy + hcen); hz0.load(sHz + hcen);
        hzj.load(sHz + hjm); hxj.load(sHx + hjm);
        const T hz_p = sHz[hcen - 1], hy_p = sHy[hcen - 1];
        ex.load(sE + oc); ey.load(sE + OWN_E + oc); ez.load(sE + 2 * OWN_E + oc);
        if (PD == 0) {
            dx.load(sA + oc); dy.load(sA + OWN_E + oc); dz.load(sA + 2 * OWN_E + oc);
            ux.load(sA + 3 * OWN_E + oc); uy.load(sA + 4 * OWN_E + oc); uz.load(sA + 5 * OWN_E + oc);
        } else if (PD == 1) dx.load(sA + oc);
        else if (PD == 2) dy.load(sA + oc);
        else if (PD == 3) dz.load(sA + oc);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int npmax = NS;
            if (GEN) {
                npmax = 0;
#pragma unroll
                for (int v = 0; v < V; ++v) npmax = max(npmax, pole_count(p, c == 0 ? mx[v] : c == 1 ? my[v] : mz[v]));
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                pol.need[c][s] = (npmax > s);
                pol.cur[c][s].load(sP + ((c * NS + s) * 2) * OWN_E + oc);
                pol.prv[c][s].load(sP + ((c * NS + s) * 2 + 1) * OWN_E + oc);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty(q));
        const unsigned smask = SRC ? src_plane_mask(p, k) : 0u;
        if (act) {
            if (PD == 4) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                    const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                    T dD[3];
                    dD[0] = -(C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]));
                    dD[1] = -(C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi));
                    dD[2] = -(C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]));
                    if (SRC && smask) {
                        T S0, S1, J;
#pragma unroll
                        for (int c = 0; c < 3; ++c) { source_parts(p, smask, c, i0 + v, j, k, set, step, S0, S1, J); dD[c] -= (S1 - S0) + J; }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int m = c == 0 ? mx[v] : c == 1 ? my[v] : mz[v];
                        T &e = c == 0 ? ex.v[v] : c == 1 ? ey.v[v] : ez.v[v];
                        if (POL) {
                            T dP = T(0);
#pragma unroll
                            for (int s = 0; s < NS; ++s) {
                                const T *cf = p.mt_coef + ((long long)m * SJ_MAX_POLES + s) * 3;
                                const T c0 = UNI ? cfu[s][0] : cf[0], c1 = UNI ? cfu[s][1] : cf[1], c2 = UNI ? cfu[s][2] : cf[2];
                                const T pcur = pol.cur[c][s].v[v];
                                const T pn = c0 * pcur + c1 * pol.prv[c][s].v[v] + c2 * e;
                                pol.prv[c][s].v[v] = pn;
                                dP += pn - pcur;
                            }
                            e += (UNI ? chi_u : p.mt_chi[m]) * (dD[c] - dP);
                        } else {
                            e += chi_u * dD[c];
                        }
                    }
                }
            } else {
                T szi = T(0), izi = T(1), szh = T(0);
                if (PD == 0 || PD == 3) { szi = p.sig[2][2 * k]; izi = p.siginv[2][2 * k]; szh = p.sig[2][2 * k + 1]; }
                const bool kin = (k >= 1 && k <= p.n[2] - 1);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int i = i0 + v;
                    const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                    const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                    const bool iin = (i >= 1 && i <= p.n[0] - 1);
                    const T sf = PD == 1 ? sxi[v] : PD == 2 ? syi : szi, isf = PD == 1 ? ixi[v] : PD == 2 ? iyi : izi;
                    T S0 = T(0), S1 = T(0), J = T(0);
                    if (i <= p.n[0] - 1 && jin && kin) {     // Ex: k-dir y, u-dir z, w-dir x
                        const T curl = C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym.v[v]);
                        if (SRC && smask) source_parts(p, smask, 0, i, j, k, set, step, S0, S1, J);
                        if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, syi, iyi, szi, izi, sxh[v], ux.v[v], mx[v], chi_u, eps_u, S0, S1, J, cfu);
                        else if (PD == 1) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, T(0), T(1), T(0), T(1), sxh[v], ux.v[v], mx[v], chi_u, eps_u, S0, S1, J, cfu);
                        else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, sf, isf, T(0), T(1), T(0), ux.v[v], mx[v], chi_u, eps_u, S0, S1, J, cfu);
                    }
                    if (j <= p.n[1] - 1 && iin && kin) {     // Ey: k-dir z, u-dir x, w-dir y
                        const T curl = C * (((hxm.v[v] - hx0.v[v]) + hz0.v[v]) - hzi);
                        if (SRC && smask) source_parts(p, smask, 1, i, j, k, set, step, S0, S1, J);
                        if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, szi, izi, sxi[v], ixi[v], syh, uy.v[v], my[v], chi_u, eps_u, S0, S1, J, cfu);
                        else if (PD == 2) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, T(0), T(1), T(0), T(1), syh, uy.v[v], my[v], chi_u, eps_u, S0, S1, J, cfu);
                        else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, sf, isf, T(0), T(1), T(0), uy.v[v], my[v], chi_u, eps_u, S0, S1, J, cfu);
                    }
                    if (k <= p.n[2] - 1 && iin && jin) {     // Ez: k-dir x, u-dir y, w-dir z
                        const T curl = C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]);
                        if (SRC && smask) source_parts(p, smask, 2, i, j, k, set, step, S0, S1, J);
                        if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, sxi[v], ixi[v], syi, iyi, szh, uz.v[v], mz[v], chi_u, eps_u, S0, S1, J, cfu);
                        else if (PD == 3) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, T(0), T(1), T(0), T(1), szh, uz.v[v], mz[v], chi_u, eps_u, S0, S1, J, cfu);
                        else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, sf, isf, T(0), T(1), T(0), uz.v[v], mz[v], chi_u, eps_u, S0, S1, J, cfu);
                    }
                }
            }
            ex.store(pE); ey.store(pE + fcs); ez.store(pE + fcs2);
            if (PD == 0 || PD == 1) dx.store(pD);
            if (PD == 0 || PD == 2) dy.store(pD + bcs);
            if (PD == 0 || PD == 3) dz.store(pD + 2 * bcs);
            if (PD == 0) { ux.store(pU); uy.store(pU + bcs); uz.store(pU + 2 * bcs); }
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (pol.need[c][s]) pol.prv[c][s].store(bprv + (3 * s + c) * pcs + xg);
        }
        hxm = hx0; hym = hy0;
        if (GEN) {
#pragma unroll
            for (int v = 0; v < V; ++v) { mx[v] = nx_[v]; my[v] = ny_[v]; mz[v] = nz_[v]; }
        }
        pE += plane; pD += bplane; pU += bplane; pm += plane; xg += plane;
    }
}

template <typename T, int NT, int NST, int NS, bool UNI, bool GENERAL, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) e_tma(const KParams<T> p, const PmlBoxSet<T> bs, const TmaPlan plan, int k_lo, int k_hi) {
    typedef EStage<NT, NS, GENERAL> ES;
    extern __shared__ unsigned char sj_tma_smem[];
    Ring<NST> r;
    ring_setup<NST>(r, sj_tma_smem, ES::BYTES, NT / 32);
    if (threadIdx.x >= NT) {
        if (threadIdx.x == NT) e_tma_produce<T, NT, NST, NS, GENERAL>(p, bs, plan, r, k_lo, k_hi);
        return;
    }
    int q = 0;
    const int n0 = plan.blk_first[blockIdx.x], n1 = plan.blk_first[blockIdx.x + 1];
    for (int n = n0; n < n1; ++n) {
        const WorkItem it = plan.items[n];
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        if (kb >= ke) continue;
        const TShape sh = plan.shape[it.shape];
        const bool src = src_in_chunk(p, kb, ke);
#define SJ_E_ITEM(PD_, SRC_) e_tma_item<T, NT, NST, NS, PD_, SRC_, UNI, GENERAL>(p, bs, it, sh, r, kb, ke, q)
        if (GENERAL) { if (src) SJ_E_ITEM(0, true); else SJ_E_ITEM(0, false); }
        else if (it.box < 0) { if (src) SJ_E_ITEM(4, true); else SJ_E_ITEM(4, false); }
        else if (it.kind == 1) { if (src) SJ_E_ITEM(1, true); else SJ_E_ITEM(1, false); }
        else if (it.kind == 2) { if (src) SJ_E_ITEM(2, true); else SJ_E_ITEM(2, false); }
        else { if (src) SJ_E_ITEM(3, true); else SJ_E_ITEM(3, false); }
#undef SJ_E_ITEM
    }
}
