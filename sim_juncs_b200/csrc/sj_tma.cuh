// sj_tma.cuh -- the TMA-staged step kernel of the FDTD hot path (sm_100a).
//
// One persistent kernel, step_tma: one block per SM pulls work items off a queue in device memory -- an xy tile x a run of z
// planes of one cell class (interior, single-sigma PML face, PML edge / corner), one half-pass and one material class -- and
// marches each item along z.  Data movement is Blackwell's: one elected producer thread issues a cp.async.bulk.tensor (TMA)
// box load per array per plane into a byte-granular ring in shared memory and arms an mbarrier with the byte count; the
// consumer warps wait on that barrier, read their cells and the i/j neighbours from the staged boxes (the curl inputs are
// loaded with a one-vector / one-row halo, so there are no shuffles, no tile-edge loads and no boundary predicates: TMA
// zero-fills outside the grid), do the arithmetic of sj_kernels.cuh and store with 128-bit st.global.  The ring is released
// load by load through a second set of mbarriers (one arrival per consumer warp), so the producer runs as many planes ahead
// as the ring holds -- also across the boundary between two items.
//
// The queue holds the H-pass items of a slab, its E-pass items, or -- the fused step, one launch per time step -- both, ordered
// as a wavefront along z with per-chunk dependency counters, so that the E-pass finds the H planes just written in the L2
// (what each piece is worth is measured in profiles/README.md, round 2).
//
// The arithmetic (curl order, UPML forms, ADE) is expression for expression that of the register kernels in
// sj_kernels.cuh, so all paths give bit-identical fields (tests/test_gpu_tma.py).
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
    unsigned long long *prof;             // -DSJ_TMA_PROF builds: [block][8] cycle counters (see tma_pass), else unused
    // fused step (step_tma): H-pass and E-pass items of one step in one queue, ordered as a wavefront along z
    const int *grp_need;                  // [chunk] consumer-warp completions that finish the H-pass items of a z chunk
    int *grp_done;                        // [chunk] running count of those completions (never reset: target = (epoch + 1) * need)
    int *epoch;                           // launches of the fused kernel so far (advanced by the last block of a launch)
};
#define SJ_EPASS_FLAG 0x200               // WorkItem::shape bit: E-pass item of a fused queue; bits 16.. = its z chunk
#ifdef SJ_TMA_PROF
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(acc, v) acc += clock64() - v
#else
#define PROF_T0(v)
#define PROF_ADD(acc, v)
#endif
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
__device__ __forceinline__ void mbar_wait_raw(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_raw(bar, parity); }
#ifdef SJ_TMA_PROF
#define MBAR_WAIT_FULL(cu, bar, parity) do { const long long t0_ = clock64(); mbar_wait_raw(bar, parity); (cu).t_full += clock64() - t0_; } while (0)
#else
#define MBAR_WAIT_FULL(cu, bar, parity) mbar_wait_raw(bar, parity)
#endif
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
// knock-out builds (scripts/build_variant.sh; wrong results, timing only): SJ_KO_STORE keeps the shared-memory reads and
// the arithmetic but never executes a global store; SJ_KO_CONSUME makes the consumers release every load untouched
#ifdef SJ_KO_STORE
#define KO_STORE_OK(p) ((p).n[0] < 0)
#else
#define KO_STORE_OK(p) true
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
    static_assert((NB & (NB - 1)) == 0, "NB is a power of two (q % NB is a mask)");
    uint32_t full0, empty0, data0;      // shared-space addresses (barriers, byte 0 of the ring)
    unsigned char *data_gen;            // generic address of byte 0
    WorkItem *mail;                     // [NB] the item that starts with load q (written by the producer before it arms full(q))
    int *vtab;                          // [NB] virtual start offset of the loads in flight -- producer's bookkeeping, in shared
                                        // memory: with the ring taking the whole SM there is no L1 left for a local array
    int cap;
    __device__ __forceinline__ uint32_t full(int q) const { return full0 + 8u * ((unsigned)q & (NB - 1)); }
    __device__ __forceinline__ uint32_t empty(int q) const { return empty0 + 8u * ((unsigned)q & (NB - 1)); }
    __device__ __forceinline__ uint32_t phase(int q) const { return ((unsigned)q / NB) & 1u; }
};
#define SJ_RING_HEADER 1152             // barriers 2 x 128 B | vtab 64 B (+ pad) | mailbox NB x 48 B

// position of the next load: the same arithmetic on the producer and on every consumer thread
struct Cursor {
    int q, head;
#ifdef SJ_TMA_PROF
    long long t_full;
#endif
    __device__ __forceinline__ int place(int len, int cap) { const int pos = (head + len > cap) ? 0 : head; head = pos + len; return pos; }
};

template <int NB>
__device__ __forceinline__ void ring_setup(Ring<NB> &r, unsigned char *smem, int cap, int n_consumer_warps) {
    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    static_assert(8 * NB <= 128 && 4 * NB <= 128 && 384 + sizeof(WorkItem) * NB <= SJ_RING_HEADER, "ring header layout");
    r.full0 = base; r.empty0 = base + 128u; r.data0 = base + SJ_RING_HEADER; r.cap = cap;
    r.data_gen = smem + (base + SJ_RING_HEADER - smem_u32(smem));
    r.vtab = reinterpret_cast<int *>(smem + (base + 256u - smem_u32(smem)));
    r.mail = reinterpret_cast<WorkItem *>(smem + (base + 384u - smem_u32(smem)));
    if (threadIdx.x == 0) {
        for (int s = 0; s < NB; ++s) { mbar_init(r.full0 + 8u * s, 1u); mbar_init(r.empty0 + 8u * s, (unsigned)n_consumer_warps); }
        mbar_fence_init();
    }
    __syncthreads();
}

// producer side: where load q goes and which earlier loads must have been released first.
// Loads are placed and released in order, so the ring is a FIFO of bytes.  Every load has a virtual start offset v (bytes
// handed out so far, the bytes skipped at a wrap-around included; position = v mod cap); a new load [v, v + len) lies on
// top of exactly the loads in flight that start before v + len - cap, i.e. a prefix of the FIFO -- one comparison against
// the oldest load's start, kept in a register, instead of a scan over the loads in flight.
template <int NB>
struct Producer {
    Cursor c;
    int oldest;                 // first load not yet known to be released
    int v, v_oldest;            // virtual end of the last load handed out; virtual start of load `oldest`
    long long t_wait, t_queue, t_start, t_issue; int n_ops;
    __device__ __forceinline__ void init() { c.q = 0; c.head = 0; oldest = 0; v = 0; v_oldest = 0; t_wait = 0; t_queue = 0; t_issue = 0; n_ops = 0; t_start = clock64(); }
    __device__ __forceinline__ void retire(const Ring<NB> &r) {
        PROF_T0(t0);
        mbar_wait(r.empty(oldest), r.phase(oldest));
        PROF_ADD(t_wait, t0);
        ++oldest;
        if (oldest < c.q) v_oldest = r.vtab[oldest & (NB - 1)];
    }
    __device__ __forceinline__ void report(const TmaPlan &plan, int n_items) {
#ifdef SJ_TMA_PROF
        if (plan.prof) {
            unsigned long long *o = plan.prof + 8 * blockIdx.x;
            atomicAdd(o + 0, (unsigned long long)(clock64() - t_start)); atomicAdd(o + 1, (unsigned long long)t_wait);
            atomicAdd(o + 2, (unsigned long long)t_queue); atomicAdd(o + 3, (unsigned long long)c.q); atomicAdd(o + 6, (unsigned long long)n_items);
            atomicAdd(o + 7, (unsigned long long)t_issue);
        }
#endif
    }
    // returns the shared-space address of the load's bytes; the caller arms r.full(c.q), issues the copies, then ++c.q
    __device__ __forceinline__ uint32_t acquire(const Ring<NB> &r, int len) {
        while (oldest <= c.q - NB) retire(r);               // barrier pair q % NB is free again
        if (c.head + len > r.cap) v += r.cap - c.head;      // the bytes skipped at the end of the ring
        const int pos = c.place(len, r.cap);
        const int lim = v + len - r.cap;
        while (oldest < c.q && v_oldest < lim) retire(r);   // every in-flight load under [pos, pos + len) has been read
        if (oldest == c.q) v_oldest = v;
        r.vtab[c.q & (NB - 1)] = v;
        v += len;
        return r.data0 + (unsigned)pos;
    }
};

#ifdef SJ_KO_CONSUME
template <int NB>
__device__ __forceinline__ void ko_item(const Ring<NB> &r, Cursor &cu, int n_loads, int len_part, int len_full) {
    for (int n = 0; n < n_loads; ++n, ++cu.q) {
        cu.place(n == 0 ? len_part : len_full, r.cap);
        MBAR_WAIT_FULL(cu, r.full(cu.q), r.phase(cu.q));
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(r.empty(cu.q));
    }
}
#endif

// z coordinate (dimension 2 of the tensor maps) of plane kl of (array a, set) for arrays stored [a][set][plane]
__device__ __forceinline__ int zcoord(int a, int n_sets, int set, int planes, int kl) { return (a * n_sets + set) * planes + kl; }

// fused step: an E-pass item of z chunk c reads the H planes of chunks c - 1 and c and overwrites E planes that the H-pass
// items of those two chunks read, so it starts when both chunks' H-pass items have finished (they sit earlier in the queue)
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_group(const TmaPlan &plan, int c_first, int c_last, int epoch, int *err) {
    const long long t0 = clock64();
    for (int c = c_first; c <= c_last; ++c) {
        const int target = (epoch + 1) * plan.grp_need[c];
        while ((int)((unsigned)ld_acquire_gpu(plan.grp_done + c) - (unsigned)target) < 0) {
            if (clock64() - t0 > 20000000000LL) { *err = 1; break; }     // ~10 s: flag it, do not hang
        }
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");
}
// consumer side of the same: when every consumer warp has issued its stores of an H-pass item (named barrier 2), one thread
// publishes them with a release at gpu scope and counts the item.  (A __threadfence per warp cost 12 % of the kernel's
// stall samples: MEMBAR.SC + CCTL.IVALL seven times per item.)
template <int NT>
__device__ __forceinline__ void group_done(int *ctr) {
    asm volatile("bar.sync 2, %0;\n" ::"n"(NT) : "memory");
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(ctr) : "memory");
}

// =====================================================================================================
// H-pass
// =====================================================================================================
// a plane load: Ex, Ey, Ez boxes with the high-side halo (sh.hs bytes each) | Hx, Hy, Hz tiles (sh.os bytes each) |
// auxiliaries: none (interior), 1 (the normal B of a face tile), 6 (edge / corner: Bx, By, Bz, Ux, Uy, Uz)
template <typename T, int NT, int NB, typename F>
__device__ __forceinline__ void h_produce_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const SlabLinks<T> &lk,
                                               const Ring<NB> &r, Producer<NB> &pr, const WorkItem &it, int k_lo, int k_hi, F &&claim_next) {
    const uint64_t pol_first = l2_evict_first(); (void)pol_first;
    bool first = true;
    const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
    if (kb >= ke) return;                             // (never in a whole-slab pass)
    const TShape sh = plan.shape[it.shape & 0xff];
    const CUtensorMap *mp = plan.maps + (it.shape & 0xff) * SJ_TMAP_PER_SHAPE;
    const uint32_t hb = (uint32_t)(sh.hp * (sh.th + 1)) * sizeof(T), ob = (uint32_t)(sh.tw * sh.th) * sizeof(T);
    const uint32_t HS = (uint32_t)sh.hs, OS = (uint32_t)sh.os;
    const bool general = it.box >= 0 && it.kind == 0;
    const int naux = general ? 6 : (it.box >= 0 ? 1 : 0);
    const int len_full = 3 * HS + (3 + naux) * OS;
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
        if (k == kb) claim_next();                    // the next item is claimed while this item's last plane is issued
        const uint32_t d = pr.acquire(r, partial ? 2 * HS : len_full), bar = r.full(pr.c.q);
        if (first) { r.mail[pr.c.q & (NB - 1)] = it; first = false; }     // ordered before the barrier arrival below (release)
        ++pr.c.q;
        const int kl = k - p.kz0 + 1;
        PROF_T0(ti);
        if (partial) {
            mbar_expect_tx(bar, 2 * hb);
            tma_load_3d(d, mp + SJ_TMAP_F_HALO, it.i0, it.j0, zcoord(0, p.n_sets, it.set, p.nzl, kl), bar);
            tma_load_3d(d + HS, mp + SJ_TMAP_F_HALO, it.i0, it.j0, zcoord(1, p.n_sets, it.set, p.nzl, kl), bar);
            PROF_ADD(pr.t_issue, ti);
            continue;
        }
        mbar_expect_tx(bar, 3 * hb + (3 + naux) * ob);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            tma_load_3d(d + c * HS, mp + SJ_TMAP_F_HALO, it.i0, it.j0, zcoord(c, p.n_sets, it.set, p.nzl, kl), bar);
        const uint32_t dh = d + 3 * HS;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            TMA_OWN(dh + c * OS, mp + SJ_TMAP_F_OWN, it.i0, it.j0, zcoord(3 + c, p.n_sets, it.set, p.nzl, kl), bar);
        if (general) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {     // B = array group 1, UB = array group 3 of the box allocation
                TMA_OWN(dh + (3 + c) * OS, mb, bi, bj, zcoord(3 + c, p.n_sets, it.set, bz, k - bk0), bar);
                TMA_OWN(dh + (6 + c) * OS, mb, bi, bj, zcoord(9 + c, p.n_sets, it.set, bz, k - bk0), bar);
            }
        } else if (naux) {
            TMA_OWN(dh + 3 * OS, mb, bi, bj, zcoord(1, p.n_sets, it.set, bz, k - bk0), bar);     // a face region stores [D_n | B_n]
        }
        PROF_ADD(pr.t_issue, ti);
    }
}

// PD: 4 interior, 1/2/3 face with normal x/y/z, 0 general
// SUB: rows per thread (tall tiles: the thread's second row lies sh.ths rows above its first; PD == 4 only)
// LATE: the load is released after the arithmetic instead of right after the shared-memory reads (fewer live registers:
// the reads can be interleaved with the arithmetic; used by the two-blocks-per-SM build)
template <typename T, int NT, int NB, int PD, bool LATE, int SUB>
__device__ __forceinline__ void h_tma_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                           const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu, int *done_ctr) {
    static_assert(SUB == 1 || PD == 4, "tall tiles: interior items only");
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NAUX = PD == 0 ? 6 : PD == 4 ? 0 : 1;
    const int HS = sh.hs / (int)sizeof(T), OS = sh.os / (int)sizeof(T);      // slot sizes in elements
    const int len_full = (3 * HS + (3 + NAUX) * OS) * (int)sizeof(T), len_part = 2 * HS * (int)sizeof(T);
    const int t = threadIdx.x, lane = t & 31;
    const int row = t / sh.nvx, vx = t - row * sh.nvx;
    const int i0 = it.i0 + vx * V, j = it.j0 + row;
    bool act[SUB];
    int hc[SUB], oc[SUB];                                       // element offsets inside a halo box / an own tile
#pragma unroll
    for (int s = 0; s < SUB; ++s) {
        const int rs = row + s * sh.ths;
        const bool in_tile = row < sh.ths && rs < sh.th;
        act[s] = in_tile && (it.j0 + rs < it.j_hi) && (i0 < it.i_hi);
        hc[s] = in_tile ? rs * sh.hp + vx * V : 0;
        oc[s] = in_tile ? rs * sh.tw + vx * V : 0;
    }
    const int hp = (row < sh.ths) ? sh.hp : 0;                  // offset of the row above inside a halo box
    const T C = p.courant;
    const long long plane = p.plane;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    const long long sub_step = (long long)sh.ths * p.pitch;
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
    if (act[0]) {
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

    Vec<T, V> ex1[SUB], ey1[SUB];
    Vec<T, V> ex0, ey0, ez0, ezj, exj, hx, hy, hz, bx, by, bz, ux, uy, uz;
    bx.zero(); by.zero(); bz.zero(); ux.zero(); uy.zero(); uz.zero();
    // the run is marched from its top plane down, so Ex, Ey of plane k + 1 are carried in registers and only one load
    // is resident at a time; first the plane above the run (Ex, Ey only)
    {
        const int pos = cu.place(len_part, r.cap);
        MBAR_WAIT_FULL(cu, r.full(cu.q), r.phase(cu.q));
        const T *d = reinterpret_cast<const T *>(r.data_gen + pos);
#pragma unroll
        for (int s = 0; s < SUB; ++s) { ex1[s].load(d + hc[s]); ey1[s].load(d + HS + hc[s]); }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty(cu.q));
        ++cu.q;
    }
    pH += (long long)(ke - 1 - kb) * plane; pB += (long long)(ke - 1 - kb) * bplane; pU += (long long)(ke - 1 - kb) * bplane;
    for (int k = ke - 1; k >= kb; --k, ++cu.q) {
        const int pos = cu.place(len_full, r.cap);
        MBAR_WAIT_FULL(cu, r.full(cu.q), r.phase(cu.q));
        const T *sEx = reinterpret_cast<const T *>(r.data_gen + pos), *sEy = sEx + HS, *sEz = sEx + 2 * HS;
        const T *sH = sEx + 3 * HS;
        if constexpr (PD == 4 && SUB == 2 && !LATE) {
            // tall interior tiles: both rows' values are read before either is computed, so that the two dependency chains
            // overlap (7 warps per SM hide little latency on their own)
            Vec<T, V> e0x[2], e0y[2], e0z[2], ejz[2], ejx[2], h_x[2], h_y[2], h_z[2];
            T nz_[2], ny_[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                e0x[s].load(sEx + hc[s]); e0y[s].load(sEy + hc[s]);
                e0z[s].load(sEz + hc[s]); ejz[s].load(sEz + hc[s] + hp); ejx[s].load(sEx + hc[s] + hp);
                nz_[s] = sEz[hc[s] + V]; ny_[s] = sEy[hc[s] + V];
                h_x[s].load(sH + oc[s]); h_y[s].load(sH + OS + oc[s]); h_z[s].load(sH + 2 * OS + oc[s]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(r.empty(cu.q));
#pragma unroll
            for (int s = 0; s < 2; ++s) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const T ezi = (v < V - 1) ? e0z[s].v[v + 1 < V ? v + 1 : v] : nz_[s];
                    const T eyi = (v < V - 1) ? e0y[s].v[v + 1 < V ? v + 1 : v] : ny_[s];
                    h_x[s].v[v] -= C * (((ejz[s].v[v] - e0z[s].v[v]) + e0y[s].v[v]) - ey1[s].v[v]);
                    h_y[s].v[v] -= C * (((ex1[s].v[v] - e0x[s].v[v]) + e0z[s].v[v]) - ezi);
                    h_z[s].v[v] -= C * (((eyi - e0y[s].v[v]) + e0x[s].v[v]) - ejx[s].v[v]);
                }
                if (act[s] && KO_STORE_OK(p)) {
                    T *pHs = pH + s * sub_step;
                    h_x[s].STORE_CS(pHs); h_y[s].STORE_CS(pHs + fcs); h_z[s].STORE_CS(pHs + fcs2);
                    if (send) {
                        T *q = lk.up.F + 3 * lk.up.fcs + (long long)it.set * lk.up.set_stride + (long long)lk.up.kl * plane + (long long)j * p.pitch + i0 + s * sub_step;
                        h_x[s].store(q); h_y[s].store(q + lk.up.fcs); h_z[s].store(q + 2 * lk.up.fcs);
                    }
                }
                ex1[s] = e0x[s]; ey1[s] = e0y[s];
            }
            pH -= plane;
            continue;
        }
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            ex0.load(sEx + hc[s]); ey0.load(sEy + hc[s]);
            ez0.load(sEz + hc[s]); ezj.load(sEz + hc[s] + hp); exj.load(sEx + hc[s] + hp);
            const T ez_n = sEz[hc[s] + V], ey_n = sEy[hc[s] + V];
            hx.load(sH + oc[s]); hy.load(sH + OS + oc[s]); hz.load(sH + 2 * OS + oc[s]);
            if (PD == 0) {
                bx.load(sH + 3 * OS + oc[s]); by.load(sH + 4 * OS + oc[s]); bz.load(sH + 5 * OS + oc[s]);
                ux.load(sH + 6 * OS + oc[s]); uy.load(sH + 7 * OS + oc[s]); uz.load(sH + 8 * OS + oc[s]);
            } else if (PD == 1) bx.load(sH + 3 * OS + oc[s]);
            else if (PD == 2) by.load(sH + 3 * OS + oc[s]);
            else if (PD == 3) bz.load(sH + 3 * OS + oc[s]);
            if (!LATE && s == SUB - 1) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
            if (act[s]) {
                if (PD == 4) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const T ezi = (v < V - 1) ? ez0.v[v + 1 < V ? v + 1 : v] : ez_n;
                        const T eyi = (v < V - 1) ? ey0.v[v + 1 < V ? v + 1 : v] : ey_n;
                        hx.v[v] -= C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1[s].v[v]);
                        hy.v[v] -= C * (((ex1[s].v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
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
                            const T curl = C * (((ezj.v[v] - ez0.v[v]) + ey0.v[v]) - ey1[s].v[v]);
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
                            const T curl = C * (((ex1[s].v[v] - ex0.v[v]) + ez0.v[v]) - ezi);
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
                if (KO_STORE_OK(p)) {
                T *pHs = pH + s * sub_step;
                hx.STORE_CS(pHs); hy.STORE_CS(pHs + fcs); hz.STORE_CS(pHs + fcs2);
                if (PD == 0 || PD == 1) bx.STORE_CS(pB);
                if (PD == 0 || PD == 2) by.STORE_CS(pB + bcs);
                if (PD == 0 || PD == 3) bz.STORE_CS(pB + 2 * bcs);
                if (PD == 0) { ux.STORE_CS(pU); uy.STORE_CS(pU + bcs); uz.STORE_CS(pU + 2 * bcs); }
                if (send) {     // the slab's top plane: the same values go straight into the upper slab's lower halo
                    T *q = lk.up.F + 3 * lk.up.fcs + (long long)it.set * lk.up.set_stride + (long long)lk.up.kl * plane + (long long)j * p.pitch + i0 + s * sub_step;
                    hx.store(q); hy.store(q + lk.up.fcs); hz.store(q + 2 * lk.up.fcs);
                }
                }
            }
            if (LATE && s == SUB - 1) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
            ex1[s] = ex0; ey1[s] = ey0;
        }
        pH -= plane; pB -= bplane; pU -= bplane;
    }
    if (send) boundary_done(lk.done, lk.n_bnd[0], lk.up.flag, (unsigned long long)*p.step + 1ull);
    if (done_ctr) group_done<NT>(done_ctr);
}

template <typename T, int NT, int NB>
__device__ __forceinline__ void h_tma_dispatch(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                               const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu, int *done_ctr) {
    if (it.box < 0) {
        if (sh.sub == 2) h_tma_item<T, NT, NB, 4, false, 2>(p, bs, lk, it, sh, r, kb, ke, cu, done_ctr);
        else h_tma_item<T, NT, NB, 4, false, 1>(p, bs, lk, it, sh, r, kb, ke, cu, done_ctr);
    }
    else if (it.kind == 0) h_tma_item<T, NT, NB, 0, false, 1>(p, bs, lk, it, sh, r, kb, ke, cu, done_ctr);
    else if (it.kind == 1) h_tma_item<T, NT, NB, 1, false, 1>(p, bs, lk, it, sh, r, kb, ke, cu, done_ctr);
    else if (it.kind == 2) h_tma_item<T, NT, NB, 2, false, 1>(p, bs, lk, it, sh, r, kb, ke, cu, done_ctr);
    else h_tma_item<T, NT, NB, 3, false, 1>(p, bs, lk, it, sh, r, kb, ke, cu, done_ctr);
}

// =====================================================================================================
// E-pass
// =====================================================================================================
// A plane load: Hx, Hy, Hz boxes with the low-side halo (origin i0 - V, j0 - 1; sh.hs bytes each) | Ex, Ey, Ez tiles
// (sh.os bytes each) | auxiliary tiles: none (interior), 1 (the normal D of a face tile), 6 (edge / corner: Dx, Dy, Dz, Ux,
// Uy, Uz) | polarisation tiles [c][s]{current, previous}.
// NS = pole slots staged (0: none); UNI: the item holds one material (it.mat) -- chi, eps and the folded ADE coefficients
// sit in registers; otherwise (mixed tiles) the material bytes are read with plain loads one plane ahead.

// material class of an item (WorkItem::pad): 0 one non-dispersive material, 1 mixed (p.n_slots slots staged), 2 / 3 one
// material with 1 / 2 poles
template <typename T>
__device__ __forceinline__ int item_slots(const KParams<T> &p, const WorkItem &it) {
    return it.pad == 0 ? 0 : it.pad == 1 ? p.n_slots : it.pad - 1;
}

template <typename T, int NT, int NB, typename F>
__device__ __forceinline__ void e_produce_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const SlabLinks<T> &lk,
                                               const Ring<NB> &r, Producer<NB> &pr, const WorkItem &it, int parity, int k_lo, int k_hi, F &&claim_next) {
    constexpr int V = 16 / (int)sizeof(T);
    const uint64_t pol_first = l2_evict_first(); (void)pol_first;
    bool first = true;
    const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
    if (kb >= ke) return;                             // (never in a whole-slab pass)
    const TShape sh = plan.shape[it.shape & 0xff];
    const CUtensorMap *mp = plan.maps + (it.shape & 0xff) * SJ_TMAP_PER_SHAPE;
    const uint32_t hb = (uint32_t)(sh.hp * (sh.th + 1)) * sizeof(T), ob = (uint32_t)(sh.tw * sh.th) * sizeof(T);
    const uint32_t HS = (uint32_t)sh.hs, OS = (uint32_t)sh.os;
    const bool general = it.box >= 0 && it.kind == 0;
    const int naux = general ? 6 : (it.box >= 0 ? 1 : 0);
    const int ns = item_slots(p, it);
    // the slab's bottom plane reads H of the plane below it: the lower slab's top plane of this step
    if ((it.shape & SJ_BND_FLAG) && lk.down.F) wait_halo(lk.flag_h, (unsigned long long)*p.step + 1ull, lk.err);
    const uint32_t off_e = 3 * HS, off_aux = off_e + 3 * OS, off_p = off_aux + naux * OS;
    const int len_full = off_p + 6 * ns * OS;
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
        if (k == ke - 1) claim_next();                // the next item is claimed while this item's last plane is issued
        const uint32_t d = pr.acquire(r, partial ? 2 * HS : len_full), bar = r.full(pr.c.q);
        if (first) { r.mail[pr.c.q & (NB - 1)] = it; first = false; }     // ordered before the barrier arrival below (release)
        ++pr.c.q;
        const int kl = k - p.kz0 + 1;
        PROF_T0(ti);
        if (partial) {
            mbar_expect_tx(bar, 2 * hb);
            tma_load_3d(d, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(3, p.n_sets, it.set, p.nzl, kl), bar);
            tma_load_3d(d + HS, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(4, p.n_sets, it.set, p.nzl, kl), bar);
            PROF_ADD(pr.t_issue, ti);
            continue;
        }
        mbar_expect_tx(bar, 3 * hb + (3 + naux + 6 * ns) * ob);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            tma_load_3d(d + c * HS, mp + SJ_TMAP_F_HALO, hi0, hj0, zcoord(3 + c, p.n_sets, it.set, p.nzl, kl), bar);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            TMA_OWN(d + off_e + c * OS, mp + SJ_TMAP_F_OWN, it.i0, it.j0, zcoord(c, p.n_sets, it.set, p.nzl, kl), bar);
        if (general) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {     // D = array group 0, UD = array group 2 of the box allocation
                TMA_OWN(d + off_aux + c * OS, mb, bi, bj, zcoord(c, p.n_sets, it.set, bz, k - bk0), bar);
                TMA_OWN(d + off_aux + (3 + c) * OS, mb, bi, bj, zcoord(6 + c, p.n_sets, it.set, bz, k - bk0), bar);
            }
        } else if (naux) {
            TMA_OWN(d + off_aux, mb, bi, bj, zcoord(0, p.n_sets, it.set, bz, k - bk0), bar);     // a face region stores [D_n | B_n]
        }
        for (int c = 0; c < 3; ++c)
            for (int s = 0; s < ns; ++s) {
                const int ac = (parity * p.n_slots + s) * 3 + c, ap = ((parity ^ 1) * p.n_slots + s) * 3 + c;
                TMA_OWN(d + off_p + ((c * ns + s) * 2) * OS, mp + SJ_TMAP_P_OWN, it.i0, it.j0, zcoord(ac, p.n_sets, it.set, p.p_nzp, kl - p.p_k0), bar);
                TMA_OWN(d + off_p + ((c * ns + s) * 2 + 1) * OS, mp + SJ_TMAP_P_OWN, it.i0, it.j0, zcoord(ap, p.n_sets, it.set, p.p_nzp, kl - p.p_k0), bar);
            }
        PROF_ADD(pr.t_issue, ti);
    }
}

// The producer thread: pulls items off the queue (the next one is fetched while the current one is issued: atomic + 48-byte
// load are ~1.5 us of latency) and issues their plane loads.  The queue holds H-pass items, E-pass items (SJ_EPASS_FLAG)
// or, for the fused step, both.
template <typename T, int NT, int NB>
__device__ __forceinline__ void tma_produce(const KParams<T> &p, const PmlBoxSet<T> &bs, const TmaPlan &plan, const SlabLinks<T> &lk,
                                            const Ring<NB> &r, int k_lo, int k_hi) {
    Producer<NB> pr; pr.init();
    int n_done = 0;
    const int parity = (int)(*p.step & 1);
    const bool fused = plan.grp_done != nullptr;
    const int epoch = fused ? *plan.epoch : 0;
    int verified = -1;                                     // fused: H-pass items of the chunks <= verified are known to be done
    // The next item is claimed (atomic + 48-byte load, ~1.5 us of latency) when the last plane of the current one is being
    // issued: late enough that a block never sits on an item it will not start for a long time (claiming a whole item ahead
    // lengthened the tail of every launch by ~8 us), early enough that the ring still holds planes to hide the latency.
    const int last = plan.n_items - 1;
    int n = atomicAdd(plan.queue, 1);
    WorkItem it = plan.items[min(n, last)];
    while (n < plan.n_items) {
        int n2 = -1;
        WorkItem it2;
        auto claim_next = [&]() { n2 = atomicAdd(plan.queue, 1); it2 = plan.items[min(n2, last)]; };
        ++n_done;
        if (it.shape & SJ_EPASS_FLAG) {
            const int chunk = it.shape >> 16;
            if (fused && chunk > verified) { PROF_T0(tq); wait_group(plan, verified + 1, chunk, epoch, lk.err); PROF_ADD(pr.t_queue, tq); verified = chunk; }
            e_produce_item<T, NT, NB>(p, bs, plan, lk, r, pr, it, parity, k_lo, k_hi, claim_next);
        } else {
            h_produce_item<T, NT, NB>(p, bs, plan, lk, r, pr, it, k_lo, k_hi, claim_next);
        }
        if (n2 < 0) claim_next();                          // (an item outside [k_lo, k_hi) issues nothing)
        n = n2; it = it2;
    }
    {   // end of queue: an empty load whose mailbox says so
        pr.acquire(r, 0);
        WorkItem e; e.box = SJ_ITEM_END;
        r.mail[pr.c.q & (NB - 1)] = e;
        mbar_arrive(r.full(pr.c.q));
        if (atomicAdd(plan.queue + 1, 1) == (int)gridDim.x - 1) { plan.queue[0] = 0; plan.queue[1] = 0; }   // ready for the next launch
        pr.report(plan, n_done);
    }
}

// SUB: rows per thread (tall tiles, PD == 4 only; see h_tma_item)
template <typename T, int NT, int NB, int NS, int PD, bool SRC, bool UNI, bool LATE, int SUB>
__device__ __forceinline__ void e_tma_item(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                           const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu, const long long step) {
    static_assert(SUB == 1 || PD == 4, "tall tiles: interior items only");
    constexpr int V = 16 / (int)sizeof(T);
    constexpr bool POL = NS > 0;
    constexpr bool GEN = POL && !UNI;             // per-cell material bytes and table look-ups
    constexpr int NS1 = NS > 0 ? NS : 1;
    constexpr int NAUX = PD == 0 ? 6 : PD == 4 ? 0 : 1;
    const int HS = sh.hs / (int)sizeof(T), OS = sh.os / (int)sizeof(T);      // slot sizes in elements
    const int len_full = (3 * HS + (3 + NAUX + 6 * NS) * OS) * (int)sizeof(T), len_part = 2 * HS * (int)sizeof(T);
    const int t = threadIdx.x, lane = t & 31;
    const int row = t / sh.nvx, vx = t - row * sh.nvx;
    const int i0 = it.i0 + vx * V, j = it.j0 + row;
    bool act[SUB];
    int hcen[SUB], oc[SUB];            // own cell inside the halo box (origin i0 - V, j0 - 1) / inside an own tile
#pragma unroll
    for (int s = 0; s < SUB; ++s) {
        const int rs = row + s * sh.ths;
        const bool in_tile = row < sh.ths && rs < sh.th;
        act[s] = in_tile && (it.j0 + rs < it.j_hi) && (i0 < it.i_hi);
        hcen[s] = in_tile ? (rs + 1) * sh.hp + (vx + 1) * V : sh.hp + V;
        oc[s] = in_tile ? rs * sh.tw + vx * V : 0;
    }
    const int hp = sh.hp;                          // the row below inside a halo box: hcen - hp
    const T C = p.courant;
    const int set = it.set;
    const int parity = (int)(step & 1);            // step: the device step counter, read once per kernel
    const long long plane = p.plane;
    const long long fcs = p.fcs, fcs2 = 2 * p.fcs;
    const long long sub_step = (long long)sh.ths * p.pitch;
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
    if (act[0]) {
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

    Vec<T, V> hxm[SUB], hym[SUB];
    Vec<T, V> hx0, hy0, hz0, hzj, hxj, ex, ey, ez, dx, dy, dz, ux, uy, uz;
    dx.zero(); dy.zero(); dz.zero(); ux.zero(); uy.zero(); uz.zero();
    unsigned char mx[SUB][V], my[SUB][V], mz[SUB][V], nx_[SUB][V], ny_[SUB][V], nz_[SUB][V];
#pragma unroll
    for (int s = 0; s < SUB; ++s) {
#pragma unroll
        for (int v = 0; v < V; ++v) mx[s][v] = my[s][v] = mz[s][v] = nx_[s][v] = ny_[s][v] = nz_[s][v] = 0;
        if (GEN && act[s]) { load_bytes<V>(pm + s * sub_step, mx[s]); load_bytes<V>(pm + mcs + s * sub_step, my[s]); load_bytes<V>(pm + mcs2 + s * sub_step, mz[s]); }
    }
    PolState<T, V, NS> pol;

    {
        const int pos = cu.place(len_part, r.cap);
        MBAR_WAIT_FULL(cu, r.full(cu.q), r.phase(cu.q));
        const T *d = reinterpret_cast<const T *>(r.data_gen + pos);
#pragma unroll
        for (int s = 0; s < SUB; ++s) { hxm[s].load(d + hcen[s]); hym[s].load(d + HS + hcen[s]); }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(r.empty(cu.q));
    ++cu.q;
    for (int k = kb; k < ke; ++k, ++cu.q) {
#pragma unroll
        for (int s = 0; s < SUB; ++s)
            if (GEN && act[s]) {
                load_bytes<V>(pm + plane + s * sub_step, nx_[s]); load_bytes<V>(pm + mcs + plane + s * sub_step, ny_[s]);
                load_bytes<V>(pm + mcs2 + plane + s * sub_step, nz_[s]);
            }
        const int pos = cu.place(len_full, r.cap);
        MBAR_WAIT_FULL(cu, r.full(cu.q), r.phase(cu.q));
        const T *sHx = reinterpret_cast<const T *>(r.data_gen + pos), *sHy = sHx + HS, *sHz = sHx + 2 * HS;
        const T *sE = sHx + 3 * HS, *sA = sE + 3 * OS, *sP = sA + NAUX * OS;
        const unsigned smask = SRC ? src_plane_mask(p, k) : 0u;
        if constexpr (PD == 4 && SUB == 2 && NS == 0 && !LATE) {
            // tall interior tiles of one non-dispersive material: both rows read first, then both computed (see h_tma_item)
            Vec<T, V> h0x[2], h0y[2], h0z[2], hjz[2], hjx[2], e_x[2], e_y[2], e_z[2];
            T pz_[2], py_[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int hc = hcen[s], o = oc[s];
                h0x[s].load(sHx + hc); h0y[s].load(sHy + hc); h0z[s].load(sHz + hc);
                hjz[s].load(sHz + hc - hp); hjx[s].load(sHx + hc - hp);
                pz_[s] = sHz[hc - 1]; py_[s] = sHy[hc - 1];
                e_x[s].load(sE + o); e_y[s].load(sE + OS + o); e_z[s].load(sE + 2 * OS + o);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(r.empty(cu.q));
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int js = j + s * sh.ths;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const T hzi = (v > 0) ? h0z[s].v[v > 0 ? v - 1 : 0] : pz_[s];
                    const T hyi = (v > 0) ? h0y[s].v[v > 0 ? v - 1 : 0] : py_[s];
                    T dD[3];
                    dD[0] = -(C * (((hjz[s].v[v] - h0z[s].v[v]) + h0y[s].v[v]) - hym[s].v[v]));
                    dD[1] = -(C * (((hxm[s].v[v] - h0x[s].v[v]) + h0z[s].v[v]) - hzi));
                    dD[2] = -(C * (((hyi - h0y[s].v[v]) + h0x[s].v[v]) - hjx[s].v[v]));
                    if (SRC && smask && act[s]) {
                        T S0, S1, J;
#pragma unroll
                        for (int c = 0; c < 3; ++c) { source_parts(p, smask, c, i0 + v, js, k, set, step, S0, S1, J); dD[c] -= (S1 - S0) + J; }
                    }
                    e_x[s].v[v] += chi_u * dD[0]; e_y[s].v[v] += chi_u * dD[1]; e_z[s].v[v] += chi_u * dD[2];
                }
                if (act[s] && KO_STORE_OK(p)) {
                    T *pEs = pE + s * sub_step;
                    e_x[s].STORE_CS(pEs); e_y[s].STORE_CS(pEs + fcs); e_z[s].STORE_CS(pEs + fcs2);
                    if (send) {
                        T *q = lk.down.F + (long long)set * lk.down.set_stride + (long long)lk.down.kl * plane + (long long)j * p.pitch + i0 + s * sub_step;
                        e_x[s].store(q); e_y[s].store(q + lk.down.fcs); e_z[s].store(q + 2 * lk.down.fcs);
                    }
                }
                hxm[s] = h0x[s]; hym[s] = h0y[s];
            }
            pE += plane; pm += plane; xg += plane;
            continue;
        }
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            const int hc = hcen[s], o = oc[s];
            hx0.load(sHx + hc); hy0.load(sHy + hc); hz0.load(sHz + hc);
            hzj.load(sHz + hc - hp); hxj.load(sHx + hc - hp);
            const T hz_p = sHz[hc - 1], hy_p = sHy[hc - 1];
            ex.load(sE + o); ey.load(sE + OS + o); ez.load(sE + 2 * OS + o);
            if (PD == 0) {
                dx.load(sA + o); dy.load(sA + OS + o); dz.load(sA + 2 * OS + o);
                ux.load(sA + 3 * OS + o); uy.load(sA + 4 * OS + o); uz.load(sA + 5 * OS + o);
            } else if (PD == 1) dx.load(sA + o);
            else if (PD == 2) dy.load(sA + o);
            else if (PD == 3) dz.load(sA + o);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int npmax = NS;
                if (GEN) {
                    npmax = 0;
#pragma unroll
                    for (int v = 0; v < V; ++v) npmax = max(npmax, pole_count(p, c == 0 ? mx[s][v] : c == 1 ? my[s][v] : mz[s][v]));
                }
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    pol.need[c][q] = (npmax > q);
                    pol.cur[c][q].load(sP + ((c * NS + q) * 2) * OS + o);
                    pol.prv[c][q].load(sP + ((c * NS + q) * 2 + 1) * OS + o);
                    // a mixed tile stages every slot of every cell; what its materials do not use is not polarisation data
                    // (outside the stored plane range the tile even aliases another array) and must read as zero
                    if (GEN && !pol.need[c][q]) { pol.cur[c][q].zero(); pol.prv[c][q].zero(); }
                }
            }
            if (!LATE && s == SUB - 1) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
            if (act[s]) {
                const int js = j + s * sh.ths;
                if (PD == 4) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const T hzi = (v > 0) ? hz0.v[v > 0 ? v - 1 : 0] : hz_p;
                        const T hyi = (v > 0) ? hy0.v[v > 0 ? v - 1 : 0] : hy_p;
                        T dD[3];
                        dD[0] = -(C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym[s].v[v]));
                        dD[1] = -(C * (((hxm[s].v[v] - hx0.v[v]) + hz0.v[v]) - hzi));
                        dD[2] = -(C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]));
                        if (SRC && smask) {
                            T S0, S1, J;
#pragma unroll
                            for (int c = 0; c < 3; ++c) { source_parts(p, smask, c, i0 + v, js, k, set, step, S0, S1, J); dD[c] -= (S1 - S0) + J; }
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int m = c == 0 ? mx[s][v] : c == 1 ? my[s][v] : mz[s][v];
                            T &e = c == 0 ? ex.v[v] : c == 1 ? ey.v[v] : ez.v[v];
                            if (POL) {
                                T dP = T(0);
#pragma unroll
                                for (int q = 0; q < NS; ++q) {
                                    const T *cf = p.mt_coef + ((long long)m * SJ_MAX_POLES + q) * 3;
                                    const T c0 = UNI ? cfu[q][0] : cf[0], c1 = UNI ? cfu[q][1] : cf[1], c2 = UNI ? cfu[q][2] : cf[2];
                                    const T pcur = pol.cur[c][q].v[v];
                                    const T pn = c0 * pcur + c1 * pol.prv[c][q].v[v] + c2 * e;
                                    pol.prv[c][q].v[v] = pn;
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
                            const T curl = C * (((hzj.v[v] - hz0.v[v]) + hy0.v[v]) - hym[s].v[v]);
                            if (SRC && smask) source_parts(p, smask, 0, i, j, k, set, step, S0, S1, J);
                            if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, syi, iyi, szi, izi, sxh[v], ux.v[v], mx[s][v], chi_u, eps_u, S0, S1, J, cfu);
                            else if (PD == 1) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, T(0), T(1), T(0), T(1), sxh[v], ux.v[v], mx[s][v], chi_u, eps_u, S0, S1, J, cfu);
                            else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 0, v, ex.v[v], dx.v[v], curl, sf, isf, T(0), T(1), T(0), ux.v[v], mx[s][v], chi_u, eps_u, S0, S1, J, cfu);
                        }
                        if (j <= p.n[1] - 1 && iin && kin) {     // Ey: k-dir z, u-dir x, w-dir y
                            const T curl = C * (((hxm[s].v[v] - hx0.v[v]) + hz0.v[v]) - hzi);
                            if (SRC && smask) source_parts(p, smask, 1, i, j, k, set, step, S0, S1, J);
                            if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, szi, izi, sxi[v], ixi[v], syh, uy.v[v], my[s][v], chi_u, eps_u, S0, S1, J, cfu);
                            else if (PD == 2) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, T(0), T(1), T(0), T(1), syh, uy.v[v], my[s][v], chi_u, eps_u, S0, S1, J, cfu);
                            else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 1, v, ey.v[v], dy.v[v], curl, sf, isf, T(0), T(1), T(0), uy.v[v], my[s][v], chi_u, eps_u, S0, S1, J, cfu);
                        }
                        if (k <= p.n[2] - 1 && iin && jin) {     // Ez: k-dir x, u-dir y, w-dir z
                            const T curl = C * (((hyi - hy0.v[v]) + hx0.v[v]) - hxj.v[v]);
                            if (SRC && smask) source_parts(p, smask, 2, i, j, k, set, step, S0, S1, J);
                            if (PD == 0) pml_e_elem<T, V, NS, 0, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, sxi[v], ixi[v], syi, iyi, szh, uz.v[v], mz[s][v], chi_u, eps_u, S0, S1, J, cfu);
                            else if (PD == 3) pml_e_elem<T, V, NS, 2, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, T(0), T(1), T(0), T(1), szh, uz.v[v], mz[s][v], chi_u, eps_u, S0, S1, J, cfu);
                            else pml_e_elem<T, V, NS, 1, SRC, UNI>(p, pol, 2, v, ez.v[v], dz.v[v], curl, sf, isf, T(0), T(1), T(0), uz.v[v], mz[s][v], chi_u, eps_u, S0, S1, J, cfu);
                        }
                    }
                }
                if (KO_STORE_OK(p)) {
                T *pEs = pE + s * sub_step;
                ex.STORE_CS(pEs); ey.STORE_CS(pEs + fcs); ez.STORE_CS(pEs + fcs2);
                if (PD == 0 || PD == 1) dx.STORE_CS(pD);
                if (PD == 0 || PD == 2) dy.STORE_CS(pD + bcs);
                if (PD == 0 || PD == 3) dz.STORE_CS(pD + 2 * bcs);
                if (PD == 0) { ux.STORE_CS(pU); uy.STORE_CS(pU + bcs); uz.STORE_CS(pU + 2 * bcs); }
#pragma unroll
                for (int q = 0; q < NS; ++q)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (pol.need[c][q]) pol.prv[c][q].STORE_CS(bprv + (3 * q + c) * pcs + (xg - psh) + s * sub_step);
                if (send) {     // the slab's bottom plane: the same values go straight into the lower slab's upper halo
                    T *q = lk.down.F + (long long)set * lk.down.set_stride + (long long)lk.down.kl * plane + (long long)j * p.pitch + i0 + s * sub_step;
                    ex.store(q); ey.store(q + lk.down.fcs); ez.store(q + 2 * lk.down.fcs);
                }
                }
            }
            if (LATE && s == SUB - 1) { __syncwarp(); if (lane == 0) mbar_arrive(r.empty(cu.q)); }
            hxm[s] = hx0; hym[s] = hy0;
            if (GEN) {
#pragma unroll
                for (int v = 0; v < V; ++v) { mx[s][v] = nx_[s][v]; my[s][v] = ny_[s][v]; mz[s][v] = nz_[s][v]; }
            }
        }
        pE += plane; pD += bplane; pU += bplane; pm += plane; xg += plane;
    }
    if (send) boundary_done(lk.done + 1, lk.n_bnd[1], lk.down.flag, (unsigned long long)step + 1ull);
}

template <typename T, int NT, int NB, int NS, bool UNI, bool LATE>
__device__ __forceinline__ void e_tma_dispatch(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                               const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu, const long long step) {
    const bool src = src_in_chunk(p, kb, ke);
#define SJ_E_ITEM(PD_, SRC_, SUB_) e_tma_item<T, NT, NB, NS, PD_, SRC_, UNI, LATE, SUB_>(p, bs, lk, it, sh, r, kb, ke, cu, step)
    if (it.box < 0) {
        // tall tiles: classes that stage at most one pole slot (the host cuts the others into half-height items)
        if (NS <= 1 && sh.sub == 2) { if (src) SJ_E_ITEM(4, true, (NS <= 1 ? 2 : 1)); else SJ_E_ITEM(4, false, (NS <= 1 ? 2 : 1)); }
        else { if (sh.sub != 1) __trap(); if (src) SJ_E_ITEM(4, true, 1); else SJ_E_ITEM(4, false, 1); }
    }
    else if (it.kind == 0) { if (src) SJ_E_ITEM(0, true, 1); else SJ_E_ITEM(0, false, 1); }
    else if (it.kind == 1) { if (src) SJ_E_ITEM(1, true, 1); else SJ_E_ITEM(1, false, 1); }
    else if (it.kind == 2) { if (src) SJ_E_ITEM(2, true, 1); else SJ_E_ITEM(2, false, 1); }
    else { if (src) SJ_E_ITEM(3, true, 1); else SJ_E_ITEM(3, false, 1); }
#undef SJ_E_ITEM
}

template <typename T, int NT, int NB>
__device__ __forceinline__ void e_tma_class(const KParams<T> &p, const PmlBoxSet<T> &bs, const SlabLinks<T> &lk, const WorkItem &it,
                                            const TShape &sh, const Ring<NB> &r, int kb, int ke, Cursor &cu, const long long step) {
    if (it.pad == 0) e_tma_dispatch<T, NT, NB, 0, true, false>(p, bs, lk, it, sh, r, kb, ke, cu, step);
    else if (it.pad == 2) e_tma_dispatch<T, NT, NB, 1, true, false>(p, bs, lk, it, sh, r, kb, ke, cu, step);
    else if (it.pad == 3) e_tma_dispatch<T, NT, NB, 2, true, false>(p, bs, lk, it, sh, r, kb, ke, cu, step);
    else if (p.n_slots <= 1) e_tma_dispatch<T, NT, NB, 1, false, false>(p, bs, lk, it, sh, r, kb, ke, cu, step);
    else e_tma_dispatch<T, NT, NB, 2, false, false>(p, bs, lk, it, sh, r, kb, ke, cu, step);
}


// The kernel of every launch: one persistent block per SM, every item class of the queue it is given -- the H-pass items of a
// slab, its E-pass items, or (fused step) both.
// Fused step, one launch per time step: the H-pass and E-pass items of the step in one queue, ordered as a wavefront along z -- H-pass
// items of chunk c + 1, then E-pass items of chunk c -- so that the E-pass finds the H planes just written, and the E planes
// the H-pass just read, in the L2 (126 MB; scripts/microbench/l2_window.cu: data survives ~50 MB of other traffic).  Per
// cell the step then moves 12 s bytes through DRAM instead of 18 s.  Dependencies are counters per z chunk (wait_group /
// group_done); every dependency of an item sits earlier in the queue and all blocks are resident, so nothing can deadlock.
template <typename T, int NT, int NB>
__global__ void __launch_bounds__(NT + 32, 1) step_tma(const KParams<T> p, const PmlBoxSet<T> bs, const TmaPlan plan, const SlabLinks<T> lk,
                                                        int cap, int k_lo, int k_hi) {
    extern __shared__ unsigned char sj_tma_smem[];
    Ring<NB> r;
    ring_setup<NB>(r, sj_tma_smem, cap, NT / 32);
    if (threadIdx.x >= NT) {
        if (threadIdx.x == NT) tma_produce<T, NT, NB>(p, bs, plan, lk, r, k_lo, k_hi);
        return;
    }
    Cursor cu; cu.q = 0; cu.head = 0;
    const long long step = *p.step;                                // (advanced by the last block to finish, see below)
#ifdef SJ_TMA_PROF
    cu.t_full = 0;
    const long long t_c0 = clock64();
#endif
    for (;;) {
        MBAR_WAIT_FULL(cu, r.full(cu.q), r.phase(cu.q));           // the first load of the next item (or the end message) is there
        const WorkItem it = r.mail[cu.q & (NB - 1)];
        if (it.box == SJ_ITEM_END) break;
        const int kb = max(it.kb, k_lo), ke = min(it.ke, k_hi);
        const TShape sh = plan.shape[it.shape & 0xff];
        if (it.shape & SJ_EPASS_FLAG) e_tma_class<T, NT, NB>(p, bs, lk, it, sh, r, kb, ke, cu, step);
        else h_tma_dispatch<T, NT, NB>(p, bs, lk, it, sh, r, kb, ke, cu, plan.grp_done ? plan.grp_done + (it.shape >> 16) : nullptr);
    }
#ifdef SJ_TMA_PROF
    if (plan.prof && (threadIdx.x & 31) == 0) {
        unsigned long long *o = plan.prof + 8 * blockIdx.x;
        if (threadIdx.x == 0) atomicAdd(o + 4, (unsigned long long)(clock64() - t_c0));
        atomicAdd(o + 5, (unsigned long long)cu.t_full);      // summed over the consumer warps
    }
#endif
    if (plan.tick || plan.epoch) {      // end of the step: every block has read the step counter and the epoch for the last time
        asm volatile("bar.sync 1, %0;\n" ::"n"(NT));      // the consumer warps of this block
        if (threadIdx.x == 0 && atomicAdd(plan.queue + 2, 1) == (int)gridDim.x - 1) {
            plan.queue[2] = 0;
            if (plan.epoch) *plan.epoch += 1;
            if (plan.tick) *plan.tick += 1;
        }
    }
}
