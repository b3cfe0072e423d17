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
// the producer runs as many planes ahead as the ring holds -- also across the boundary between two items -- whatever the occupancy.
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
    const WorkItem *items;                // sorted: heaviest first, fine-grained ones last
    int n_items;
    int *queue;                           // [0] next item to hand out, [1] producers that have drained the queue,
                                          // [2] blocks whose consumers have finished
    long long *tick;                      // E-pass kernel of a full step: the last block to finish advances the step counter
};
#define SJ_ITEM_END (-99)                 // WorkItem::box of the end-of-queue message

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

// Experiment kept behind SJ_TMA_L2_HINTS (off): L2 evict-first priority for what is read exactly once per pass (own-cell
// tiles, auxiliaries, polarisation) and streaming stores, so that the halo rows of the curl inputs would survive in the L2.
// Measured on the bench workload: L2 read hit rate unchanged (11.8 %), DRAM bytes unchanged, step 0.522 -> 0.560 ms.
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol) : "memory");
}
#ifdef SJ_TMA_L2_HINTS
#define TMA_OWN(dst, map, c0, c1, c2, bar) tma_load_3d_hint(dst, map, c0, c1, c2, bar, pol_first)
#define STORE_CS store_cs
#else
#define TMA_OWN(dst, map, c0, c1, c2, bar) tma_load_3d(dst, map, c0, c1, c2, bar)
#define STORE_CS store
#endif

// ---- slab exchange helpers -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// producer side: wait until the neighbour's plane for this step is in my halo, then order the TMA reads after it
__device__ __forceinline__ void wait_halo(const unsigned long long *flag, unsigned long long need, int *err) {
    if (!flag) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < need) {
        if (clock64() - t0 > 20000000000LL) { *err = 1; break; }     // ~10 s: the neighbour is gone; flag it, do not hang
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");
}
// consumer side, after a boundary item's remote stores: the warp that completes the plane publishes it
__device__ __forceinline__ void boundary_done(unsigned int *done, int n_total, unsigned long long *peer_flag, unsigned long long value) {
    __threadfence_system();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const unsigned int old = atomicAdd(done, 1u);
        if ((int)old == n_total - 1) {
            *done = 0u;
            __threadfence_system();
            st_release_sys(peer_flag, value);
        }
    }
}

// ---- the staging ring ------------------------------------------------------------------------------------
// One ring of `cap` bytes per block holds the plane loads of every item class back to back: load number q (counted per
// block, in the order both sides derive from the item list) takes `len` contiguous bytes at the head, or at offset 0 if
// it does not fit before the end.  Barrier pair q % NB belongs to load q: `full` completes when the TMA bytes have
// landed (one producer arrival + expect_tx), `empty` when every consumer warp has read them.  A light class (7 tiles per
// plane) therefore gets many planes in flight and a heavy one (19 tiles) two or three, from the same shared memory.
template <int NB>
struct Ring {
    uint32_t full0, empty0, data0;      // shared-space addresses (barriers, byte 0 of the ring)
    unsigned char *data_gen;            // generic address of byte 0
    WorkItem *mail;                     // [NB] the item that starts with load q (written by the producer before it arms full(q))
    int2 *meta;                         // [NB] (position, length) of the loads in flight -- producer's bookkeeping, in shared
                                        // memory: with the ring taking the whole SM there is no L1 left for a local array
    int cap;
    __device__ __forceinline__ uint32_t full(int q) const { return full0 + 8u * (unsigned)(q % NB); }
    __device__ __forceinline__ uint32_t empty(int q) const { return empty0 + 8u * (unsigned)(q % NB); }
    __device__ __forceinline__ uint32_t phase(int q) const { return (unsigned)(q / NB) & 1u; }
};

// position of the next load: the same arithmetic on the producer and on every consumer thread
struct Cursor {
    int q, head;
    __device__ __forceinline__ int place(int len, int cap) { const int pos = (head + len > cap) ? 0 : head; head = pos + len; return pos; }
};

template <int NT> struct Slots { static constexpr int OWN = NT * 16, HALO = NT * 20; };

template <int NB>
__device__ __forceinline__ void ring_setup(Ring<NB> &r, unsigned char *smem, int cap, int n_consumer_warps) {
    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    static_assert(16 * NB <= 256 && 8 * NB <= 128 && sizeof(WorkItem) * NB <= 640, "ring header layout");
    r.full0 = base; r.empty0 = base + 8u * NB; r.data0 = base + 1024u; r.cap = cap;
    r.data_gen = smem + (base + 1024u - smem_u32(smem));
    r.meta = reinterpret_cast<int2 *>(smem + (base + 256u - smem_u32(smem)));
    r.mail = reinterpret_cast<WorkItem *>(smem + (base + 384u - smem_u32(smem)));
    if (threadIdx.x == 0) {
        for (int s = 0; s < NB; ++s) { mbar_init(r.full0 + 8u * s, 1u); mbar_init(r.empty0 + 8u * s, (unsigned)n_consumer_warps); }
        mbar_fence_init();
    }
    __syncthreads();
}

// producer side: where load q goes and which earlier loads must have been released first
template <int NB>
struct Producer {
    Cursor c;
    int oldest;
    __device__ __forceinline__ void init() { c.q = 0; c.head = 0; oldest = 0; }
    __device__ __forceinline__ void wait_released(const Ring<NB> &r, int upto) {
        while (oldest <= upto) { mbar_wait(r.empty(oldest), r.phase(oldest)); ++oldest; }
    }
    // returns the shared-space address of the load's bytes; the caller arms r.full(c.q), issues the copies, then ++c.q
    __device__ __forceinline__ uint32_t acquire(const Ring<NB> &r, int len) {
        wait_released(r, c.q - NB);                         // barrier pair q % NB is free again
        const int pos = c.place(len, r.cap);
        int need = oldest - 1;
        for (int i = oldest; i < c.q; ++i) {
            const int2 m = r.meta[i % NB];
            if (m.x < pos + len && pos < m.x + m.y) need = i;
        }
        wait_released(r, need);                             // every in-flight load under [pos, pos + len) has been read
        r.meta[c.q % NB] = make_int2(pos, len);
        return r.data0 + (unsigned)pos;
    }
};

// z coordinate (dimension 2 of the tensor maps) of plane kl of (array a, set) for arrays stored [a][set][plane]
__device__ __forceinline__ int zcoord(int a, int n_sets, int set, int planes, int kl) { return (a * n_sets + set) * planes + kl; }

// =====================================================================================================
// H-pass
// =====================================================================================================
// aux slots: class A (GENERAL = false): 1 (the normal B of a face tile; unused by interior tiles);
//            GENERAL: 6 (Bx, By, Bz, Ux, Uy, Uz)
template <typename T, int NT, int NB>
__device__ __forceinline__ void h_tma_produce(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const SlabLinks<T> &lk,
                                              const Ring<NB> &r, int k_lo, int k_hi) {
    typedef Slots<NT> SL;
    Producer<NB> pr; pr.init();
    const uint64_t pol_first = l2_evict_first(); (void)pol_first;
    for (;;) {
        const int n = atomicAdd(plan.queue, 1);
        if (n >= plan.n_items) break;
        const WorkItem it = plan.items[n];
        bool first = true;
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        if (kb >= ke) continue;                           // (never in a whole-slab pass)
        const TShape sh = plan.shape[it.shape & 0xff];
        const CUtensorMap *mp = plan.maps + (it.shape & 0xff) * SJ_TMAP_PER_SHAPE;
        const uint32_t hb = (uint32_t)(sh.hp * (sh.th + 1)) * sizeof(T), ob = (uint32_t)(sh.tw * sh.th) * sizeof(T);
        const bool general = it.box >= 0 && it.kind == 0;
        const int naux = general ? 6 : (it.box >= 0 ? 1 : 0);
        const int len_full = 3 * SL::HALO + (3 + (general ? 6 : 1)) * SL::OWN;
        // the slab's top plane reads E of the plane above it: the upper slab's bottom plane of the previous step
        if ((it.shape & SJ_BND_FLAG) && lk.up.F) wait_halo(lk.flag_e, (unsigned long long)*p.step, lk.err);
        int bi = 0, bj = 0, bz = 1, bk0 = 0;
        const CUtensorMap *mb = mp;
        if (it.box >= 0) {
            const PmlBox<T> &b = bs.b[it.box];
            bi = it.i0 - b.lo[0]; bj = it.j0 - b.lo[1]; bz = b.hi[2] - b.lo[2]; bk0 = b.lo[2];
            mb = mp + SJ_TMAP_BOX0 + it.box;
        }
        for (int k = ke; k >= kb; --k) {                  // top down: plane k + 1 is then always the previous load
            const bool partial = (k == ke);               // the plane above the run: Ex, Ey only
            const uint32_t d = pr.acquire(r, partial ? 2 * SL::HALO : len_full), bar = r.full(pr.c.q);
            if (first) { r.mail[pr.c.q % NB] = it; first = false; }     // ordered before the barrier arrival below (release)
            ++pr.c.q;
            const int kl = k - p.kz0 + 1;
            if (partial) {
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
                TMA_OWN(d + 3 * SL::HALO + c * SL::OWN, mp + SJ_TMAP_F_OWN, it.i0, it.j0, zcoord(3 + c, p.n_sets, it.set, p.nzl, kl), bar);
            if (general) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {     // B = array group 1, UB = array group 3 of the box allocation
                    TMA_OWN(d + 3 * SL::HALO + (3 + c) * SL::OWN, mb, bi, bj, zcoord(3 + c, p.n_sets, it.set, bz, k - bk0), bar);
                    TMA_OWN(d + 3 * SL::HALO + (6 + c) * SL::OWN, mb, bi, bj, zcoord(9 + c, p.n_sets, it.set, bz, k - bk0), bar);
                }
            } else if (naux) {
                TMA_OWN(d + 3 * SL::HALO + 3 * SL::OWN, mb, bi, bj, zcoord(1, p.n_sets, it.set, bz, k - bk0), bar);     // a face region stores [D_n | B_n]
            }
        }
    }
    {   // end of queue: an empty load whose mailbox says so
        pr.acquire(r, 0);
        WorkItem e; e.box = SJ_ITEM_END;
        r.mail[pr.c.q % NB] = e;
        mbar_arrive(r.full(pr.c.q));
        if (atomicAdd(plan.queue + 1, 1) == (int)gridDim.x - 1) { plan.queue[0] = 0; plan.queue[1] = 0; }   // ready for the next launch
    }
}

// PD: 4 interior, 1/2/3 face with normal x/y/z, 0 general
// LATE: the load is released after the arithmetic instead of right after the shared-memory reads (fewer live registers:
// the reads can be interleaved with the arithmetic; used by the two-blocks-per-SM build)
template <typename T, int NT, int NB, int PD, bool LATE>
__device__ __forceinline__ void h_tma_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                           const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu) {
    typedef Slots<NT> SL;
    constexpr int LEN_FULL = 3 * SL::HALO + (3 + (PD == 0 ? 6 : 1)) * SL::OWN, LEN_PART = 2 * SL::HALO;
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
    const bool send = (it.shape & SJ_BND_FLAG) && lk.up.F != nullptr;      // block-uniform

    Vec<T, V> ex0, ey0, ex1, ey1, ez0, ezj, exj, hx, hy, hz, bx, by, bz, ux, uy, uz;
    bx.zero(); by.zero(); bz.zero(); ux.zero(); uy.zero(); uz.zero();
    // the run is marched from its top plane down, so Ex, Ey of plane k + 1 are carried in registers and only one load
    // is resident at a time; first the plane above the run (Ex, Ey only)
    {
        const int pos = cu.place(LEN_PART, r.cap);
        mbar_wait(r.full(cu.q), r.phase(cu.q));
        const unsigned char *d = r.data_gen + pos;
        ex1.load(reinterpret_cast<const T *>(d) + hc); ey1.load(reinterpret_cast<const T *>(d + SL::HALO) + hc);
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty(cu.q));
        ++cu.q;
    }
    pH += (long long)(ke - 1 - kb) * plane; pB += (long long)(ke - 1 - kb) * bplane; pU += (long long)(ke - 1 - kb) * bplane;
    for (int k = ke - 1; k >= kb; --k, ++cu.q) {
        const int pos = cu.place(LEN_FULL, r.cap);
        mbar_wait(r.full(cu.q), r.phase(cu.q));
        const unsigned char *d0 = r.data_gen + pos;
        const T *sEx = reinterpret_cast<const T *>(d0), *sEy = reinterpret_cast<const T *>(d0 + SL::HALO), *sEz = reinterpret_cast<const T *>(d0 + 2 * SL::HALO);
        const T *sH = reinterpret_cast<const T *>(d0 + 3 * SL::HALO);
        constexpr int OWN_E = SL::OWN / (int)sizeof(T);
        ex0.load(sEx + hc); ey0.load(sEy + hc);
        ez0.load(sEz + hc); ezj.load(sEz + hj); exj.load(sEx + hj);
        const T ez_n = sEz[hc + V], ey_n = sEy[hc + V];
        hx.load(sH + oc); hy.load(sH + OWN_E + oc); hz.load(sH + 2 * OWN_E + oc);
        if (PD == 0) {
            bx.load(sH + 3 * OWN_E + oc); by.load(sH + 4 * OWN_E + oc); bz.load(sH + 5 * OWN_E + oc);
            ux.load(sH + 6 * OWN_E + oc); uy.load(sH + 7 * OWN_E + oc); uz.load(sH + 8 * OWN_E + oc);
        } else if (PD == 1) bx.load(sH + 3 * OWN_E + oc);
        else if (PD == 2) by.load(sH + 3 * OWN_E + oc);
        else if (PD == 3) bz.load(sH + 3 * OWN_E + oc);
        if (!LATE) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
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
            hx.STORE_CS(pH); hy.STORE_CS(pH + fcs); hz.STORE_CS(pH + fcs2);
            if (PD == 0 || PD == 1) bx.STORE_CS(pB);
            if (PD == 0 || PD == 2) by.STORE_CS(pB + bcs);
            if (PD == 0 || PD == 3) bz.STORE_CS(pB + 2 * bcs);
            if (PD == 0) { ux.STORE_CS(pU); uy.STORE_CS(pU + bcs); uz.STORE_CS(pU + 2 * bcs); }
            if (send) {     // the slab's top plane: the same values go straight into the upper slab's lower halo
                T *q = lk.up.F + 3 * lk.up.fcs + (long long)it.set * lk.up.set_stride + (long long)lk.up.kl * plane + (long long)j * p.pitch + i0;
                hx.store(q); hy.store(q + lk.up.fcs); hz.store(q + 2 * lk.up.fcs);
            }
        }
        if (LATE) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
        ex1 = ex0; ey1 = ey0;
        pH -= plane; pB -= bplane; pU -= bplane;
    }
    if (send) boundary_done(lk.done, lk.n_bnd[0], lk.up.flag, (unsigned long long)*p.step + 1ull);
}

template <typename T, int NT, int NB, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) h_tma(const KParams<T> p, const PmlBoxSet<T> bs, const TmaPlan plan, const SlabLinks<T> lk,
                                                         int cap, int k_lo, int k_hi) {
    extern __shared__ unsigned char sj_tma_smem[];
    Ring<NB> r;
    ring_setup<NB>(r, sj_tma_smem, cap, NT / 32);
    if (threadIdx.x >= NT) {
        if (threadIdx.x == NT) h_tma_produce<T, NT, NB>(p, bs, plan, lk, r, k_lo, k_hi);
        return;
    }
    Cursor cu; cu.q = 0; cu.head = 0;
    for (;;) {
        mbar_wait(r.full(cu.q), r.phase(cu.q));           // the first load of the next item (or the end message) is there
        const WorkItem it = r.mail[cu.q % NB];
        if (it.box == SJ_ITEM_END) break;
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        const TShape sh = plan.shape[it.shape & 0xff];
        if (it.box < 0) h_tma_item<T, NT, NB, 4, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else if (it.kind == 0) h_tma_item<T, NT, NB, 0, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else if (it.kind == 1) h_tma_item<T, NT, NB, 1, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else if (it.kind == 2) h_tma_item<T, NT, NB, 2, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else h_tma_item<T, NT, NB, 3, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
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

// material class of an item (WorkItem::pad): 0 one non-dispersive material, 1 mixed (p.n_slots slots staged), 2 / 3 one
// material with 1 / 2 poles
template <typename T>
__device__ __forceinline__ int item_slots(const KParams<T> &p, const WorkItem &it) {
    return it.pad == 0 ? 0 : it.pad == 1 ? p.n_slots : it.pad - 1;
}

template <typename T, int NT, int NB>
__device__ __forceinline__ void e_tma_produce(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const SlabLinks<T> &lk,
                                              const Ring<NB> &r, int k_lo, int k_hi) {
    typedef Slots<NT> SL;
    constexpr int V = 16 / (int)sizeof(T);
    Producer<NB> pr; pr.init();
    const uint64_t pol_first = l2_evict_first(); (void)pol_first;
    const int parity = (int)(*p.step & 1);
    for (;;) {
        const int n = atomicAdd(plan.queue, 1);
        if (n >= plan.n_items) break;
        const WorkItem it = plan.items[n];
        bool first = true;
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        if (kb >= ke) continue;                           // (never in a whole-slab pass)
        const TShape sh = plan.shape[it.shape & 0xff];
        const CUtensorMap *mp = plan.maps + (it.shape & 0xff) * SJ_TMAP_PER_SHAPE;
        const uint32_t hb = (uint32_t)(sh.hp * (sh.th + 1)) * sizeof(T), ob = (uint32_t)(sh.tw * sh.th) * sizeof(T);
        const bool general = it.box >= 0 && it.kind == 0;
        const int naux = general ? 6 : (it.box >= 0 ? 1 : 0);
        const int ns = item_slots(p, it);
        // the slab's bottom plane reads H of the plane below it: the lower slab's top plane of this step
        if ((it.shape & SJ_BND_FLAG) && lk.down.F) wait_halo(lk.flag_h, (unsigned long long)*p.step + 1ull, lk.err);
        // same layout as EStage<NT, ns, general>
        const int off_e = 3 * SL::HALO, off_aux = off_e + 3 * SL::OWN, off_p = off_aux + (general ? 6 : 1) * SL::OWN;
        const int len_full = off_p + 6 * ns * SL::OWN;
        int bi = 0, bj = 0, bz = 1, bk0 = 0;
        const CUtensorMap *mb = mp;
        if (it.box >= 0) {
            const PmlBox<T> &b = bs.b[it.box];
            bi = it.i0 - b.lo[0]; bj = it.j0 - b.lo[1]; bz = b.hi[2] - b.lo[2]; bk0 = b.lo[2];
            mb = mp + SJ_TMAP_BOX0 + it.box;
        }
        const int hi0 = it.i0 - V, hj0 = it.j0 - 1;
        for (int k = kb - 1; k < ke; ++k) {
            const bool partial = (k == kb - 1);           // the plane below the run: Hx, Hy only
            const uint32_t d = pr.acquire(r, partial ? 2 * SL::HALO : len_full), bar = r.full(pr.c.q);
            if (first) { r.mail[pr.c.q % NB] = it; first = false; }     // ordered before the barrier arrival below (release)
            ++pr.c.q;
            const int kl = k - p.kz0 + 1;
            if (partial) {
                mbar_expect_tx(bar, 2 * hb);
                tma_load_3d(d, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(3, p.n_sets, it.set, p.nzl, kl), bar);
                tma_load_3d(d + SL::HALO, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(4, p.n_sets, it.set, p.nzl, kl), bar);
                continue;
            }
            mbar_expect_tx(bar, 3 * hb + (3 + naux + 6 * ns) * ob);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                tma_load_3d(d + c * SL::HALO, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(3 + c, p.n_sets, it.set, p.nzl, kl), bar);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                TMA_OWN(d + off_e + c * SL::OWN, mp + SJ_TMAP_F_OWN, it.i0, it.j0, zcoord(c, p.n_sets, it.set, p.nzl, kl), bar);
            if (general) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {     // D = array group 0, UD = array group 2 of the box allocation
                    TMA_OWN(d + off_aux + c * SL::OWN, mb, bi, bj, zcoord(c, p.n_sets, it.set, bz, k - bk0), bar);
                    TMA_OWN(d + off_aux + (3 + c) * SL::OWN, mb, bi, bj, zcoord(6 + c, p.n_sets, it.set, bz, k - bk0), bar);
                }
            } else if (naux) {
                TMA_OWN(d + off_aux, mb, bi, bj, zcoord(0, p.n_sets, it.set, bz, k - bk0), bar);     // a face region stores [D_n | B_n]
            }
            for (int c = 0; c < 3; ++c)
                for (int s = 0; s < ns; ++s) {
                    const int ac = (parity * p.n_slots + s) * 3 + c, ap = ((parity ^ 1) * p.n_slots + s) * 3 + c;
                    TMA_OWN(d + off_p + ((c * ns + s) * 2) * SL::OWN, mp + SJ_TMAP_P_OWN, it.i0, it.j0, zcoord(ac, p.n_sets, it.set, p.p_nzp, kl - p.p_k0), bar);
                    TMA_OWN(d + off_p + ((c * ns + s) * 2 + 1) * SL::OWN, mp + SJ_TMAP_P_OWN, it.i0, it.j0, zcoord(ap, p.n_sets, it.set, p.p_nzp, kl - p.p_k0), bar);
                }
        }
    }
    {   // end of queue: an empty load whose mailbox says so
        pr.acquire(r, 0);
        WorkItem e; e.box = SJ_ITEM_END;
        r.mail[pr.c.q % NB] = e;
        mbar_arrive(r.full(pr.c.q));
        if (atomicAdd(plan.queue + 1, 1) == (int)gridDim.x - 1) { plan.queue[0] = 0; plan.queue[1] = 0; }   // ready for the next launch
    }
}

template <typename T, int NT, int NB, int NS, int PD, bool SRC, bool UNI, bool LATE>
__device__ __forceinline__ void e_tma_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                           const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu) {
    typedef Slots<NT> SL;
    typedef EStage<NT, NS, PD == 0> ES;
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
    const long long pcs = p.p_comp_stride, psh = pol_shift(p, set);
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
    const bool send = (it.shape & SJ_BND_FLAG) && lk.down.F != nullptr;    // block-uniform
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

    {
        const int pos = cu.place(2 * SL::HALO, r.cap);
        mbar_wait(r.full(cu.q), r.phase(cu.q));
        const unsigned char *d = r.data_gen + pos;
        hxm.load(reinterpret_cast<const T *>(d) + hcen); hym.load(reinterpret_cast<const T *>(d + SL::HALO) + hcen);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(r.empty(cu.q));
    ++cu.q;
    for (int k = kb; k < ke; ++k, ++cu.q) {
        if (GEN && act) { load_bytes<V>(pm + plane, nx_); load_bytes<V>(pm + mcs + plane, ny_); load_bytes<V>(pm + mcs2 + plane, nz_); }
        const int pos = cu.place(ES::BYTES, r.cap);
        mbar_wait(r.full(cu.q), r.phase(cu.q));
        const unsigned char *d0 = r.data_gen + pos;
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
                // a mixed tile stages every slot of every cell; what its materials do not use is not polarisation data
                // (outside the stored plane range the tile even aliases another array) and must read as zero
                if (GEN && !pol.need[c][s]) { pol.cur[c][s].zero(); pol.prv[c][s].zero(); }
            }
        }
        if (!LATE) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
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
            ex.STORE_CS(pE); ey.STORE_CS(pE + fcs); ez.STORE_CS(pE + fcs2);
            if (PD == 0 || PD == 1) dx.STORE_CS(pD);
            if (PD == 0 || PD == 2) dy.STORE_CS(pD + bcs);
            if (PD == 0 || PD == 3) dz.STORE_CS(pD + 2 * bcs);
            if (PD == 0) { ux.STORE_CS(pU); uy.STORE_CS(pU + bcs); uz.STORE_CS(pU + 2 * bcs); }
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (pol.need[c][s]) pol.prv[c][s].STORE_CS(bprv + (3 * s + c) * pcs + (xg - psh));
            if (send) {     // the slab's bottom plane: the same values go straight into the lower slab's upper halo
                T *q = lk.down.F + (long long)set * lk.down.set_stride + (long long)lk.down.kl * plane + (long long)j * p.pitch + i0;
                ex.store(q); ey.store(q + lk.down.fcs); ez.store(q + 2 * lk.down.fcs);
            }
        }
        if (LATE) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
        hxm = hx0; hym = hy0;
        if (GEN) {
#pragma unroll
            for (int v = 0; v < V; ++v) { mx[v] = nx_[v]; my[v] = ny_[v]; mz[v] = nz_[v]; }
        }
        pE += plane; pD += bplane; pU += bplane; pm += plane; xg += plane;
    }
    if (send) boundary_done(lk.done + 1, lk.n_bnd[1], lk.down.flag, (unsigned long long)step + 1ull);
}

template <typename T, int NT, int NB, int NS, bool UNI, bool LATE>
__device__ __forceinline__ void e_tma_dispatch(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                               const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu) {
    const bool src = src_in_chunk(p, kb, ke);
#define SJ_E_ITEM(PD_, SRC_) e_tma_item<T, NT, NB, NS, PD_, SRC_, UNI, LATE>(p, bs, lk, it, sh, r, kb, ke, cu)
    if (it.box < 0) { if (src) SJ_E_ITEM(4, true); else SJ_E_ITEM(4, false); }
    else if (it.kind == 0) { if (src) SJ_E_ITEM(0, true); else SJ_E_ITEM(0, false); }
    else if (it.kind == 1) { if (src) SJ_E_ITEM(1, true); else SJ_E_ITEM(1, false); }
    else if (it.kind == 2) { if (src) SJ_E_ITEM(2, true); else SJ_E_ITEM(2, false); }
    else { if (src) SJ_E_ITEM(3, true); else SJ_E_ITEM(3, false); }
#undef SJ_E_ITEM
}

// One launch per E-pass: every item class (interior / face / edge tiles x material class) through the same ring.
template <typename T, int NT, int NB, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) e_tma(const KParams<T> p, const PmlBoxSet<T> bs, const TmaPlan plan, const SlabLinks<T> lk,
                                                         int cap, int k_lo, int k_hi) {
    extern __shared__ unsigned char sj_tma_smem[];
    Ring<NB> r;
    ring_setup<NB>(r, sj_tma_smem, cap, NT / 32);
    if (threadIdx.x >= NT) {
        if (threadIdx.x == NT) e_tma_produce<T, NT, NB>(p, bs, plan, lk, r, k_lo, k_hi);
        return;
    }
    Cursor cu; cu.q = 0; cu.head = 0;
    for (;;) {
        mbar_wait(r.full(cu.q), r.phase(cu.q));           // the first load of the next item (or the end message) is there
        const WorkItem it = r.mail[cu.q % NB];
        if (it.box == SJ_ITEM_END) break;
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        const TShape sh = plan.shape[it.shape & 0xff];
        if (it.pad == 0) e_tma_dispatch<T, NT, NB, 0, true, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else if (it.pad == 2) e_tma_dispatch<T, NT, NB, 1, true, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else if (it.pad == 3) e_tma_dispatch<T, NT, NB, 2, true, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else if (p.n_slots <= 1) e_tma_dispatch<T, NT, NB, 1, false, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
        else e_tma_dispatch<T, NT, NB, 2, false, MINB == 2>(p, bs, lk, it, sh, r, kb, ke, cu);
    }
    if (plan.tick) {        // end of the step: every block has read the step counter for the last time
        asm volatile("bar.sync 1, %0;\n" ::"n"(NT));      // the consumer warps of this block
        if (threadIdx.x == 0 && atomicAdd(plan.queue + 2, 1) == (int)gridDim.x - 1) { plan.queue[2] = 0; *plan.tick += 1; }
    }
}
