// sj_tma_launch.cuh -- launch logic of the TMA-staged kernels for one precision (instantiated by sj_tma_f64.cu / _f32.cu)
#pragma once
#include "sj_launch.cuh"
#include "sj_tma.cuh"

template <typename T>
static void tma_plan(const sj_sim *s, const TmaList &l, TmaPlan &plan) {
    plan.maps = (const CUtensorMap *)s->tma.maps;
    for (int i = 0; i < SJ_TMA_MAX_SHAPES; ++i) plan.shape[i] = s->tma.shapes[i];
    plan.items = l.items; plan.n_items = l.n_items; plan.queue = l.first; plan.tick = nullptr; plan.prof = nullptr; plan.grp_need = nullptr; plan.grp_done = nullptr; plan.epoch = nullptr;
}

static int tma_env(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

#ifdef SJ_TMA_PROF
// cycle counters of the producer / consumer sides, accumulated over launches and printed once (SJ_NO_GRAPH=1 runs)
static unsigned long long *tma_prof_buf(int which, int grid, cudaStream_t st) {
    static unsigned long long *buf[2] = {nullptr, nullptr};
    static int calls[2] = {0, 0};
    const int skip = tma_env("SJ_TMA_PROF_SKIP", 50), dump = tma_env("SJ_TMA_PROF_DUMP", 250);
    if (!buf[which]) { cudaMalloc(&buf[which], 8 * 1024 * sizeof(unsigned long long)); cudaMemset(buf[which], 0, 8 * 1024 * sizeof(unsigned long long)); }
    const int c = calls[which]++;
    if (c == skip) { cudaStreamSynchronize(st); cudaMemset(buf[which], 0, 8 * 1024 * sizeof(unsigned long long)); }
    if (c == dump) {
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h(8 * grid);
        cudaMemcpy(h.data(), buf[which], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        const double nl = dump - skip;
        double m[8] = {0}, mx[8] = {0}, mn[8];
        for (int q = 0; q < 8; ++q) mn[q] = 1e300;
        for (int b = 0; b < grid; ++b) for (int q = 0; q < 8; ++q) { const double v = h[8 * b + q] / nl; m[q] += v / grid; mx[q] = v > mx[q] ? v : mx[q]; mn[q] = v < mn[q] ? v : mn[q]; }
        fprintf(stderr, "[tma prof %s] per launch and block (mean [min..max]): producer cycles %.0f [%.0f..%.0f], waiting for empty %.0f [%.0f..%.0f] (%.1f %%), "
                "queue+item fetch %.0f (%.1f %%), expect_tx + TMA issue %.0f (%.1f %%), loads %.1f, items %.1f; consumer cycles %.0f [%.0f..%.0f], waiting for full (mean over 7 warps) %.0f [%.0f..%.0f] (%.1f %%)\n",
                which ? "E" : "H", m[0], mn[0], mx[0], m[1], mn[1], mx[1], 100 * m[1] / m[0], m[2], 100 * m[2] / m[0], m[7], 100 * m[7] / m[0], m[3], m[6], m[4], mn[4], mx[4],
                m[5] / 7, mn[5] / 7, mx[5] / 7, 100 * m[5] / 7 / m[4]);
    }
    return (c >= skip && c < dump) || c < skip ? buf[which] : nullptr;
}
#endif


// one persistent kernel per half-pass: one block per SM, the whole shared memory as its staging ring
template <typename T>
static int tma_pass(sj_sim *s, int which, cudaStream_t st, bool tick) {
    KParams<T> p; fill_params(s, p);
    PmlBoxSet<T> bs; memset(&bs, 0, sizeof bs);
    for (size_t bi = 0; bi < s->boxes.size(); ++bi) fill_box(s->boxes[bi], bs.b[bi], s->g.n_sets);
    constexpr int NT = SJ_TMA_NT, NB = 16;
    // which: 0 the H-pass, 1 the E-pass, 2 the whole step (fused wavefront list)
    const TmaList &l = which == 0 ? s->tma.h[0] : which == 1 ? s->tma.e[0][0] : s->tma.f;
    if (!l.n_items) return 0;
    TmaPlan plan; tma_plan<T>(s, l, plan);
    if (which >= 1 && tick) plan.tick = s->step_dev;
    if (which == 2) { plan.grp_need = s->tma.grp; plan.grp_done = s->tma.grp + s->tma.n_chunks; plan.epoch = s->tma.grp + 2 * s->tma.n_chunks; }
#ifdef SJ_TMA_PROF
    plan.prof = tma_prof_buf(which == 2 ? 1 : which, l.grid, st);
#endif
    SlabLinks<T> lk; memset(&lk, 0, sizeof lk);
    if (s->peer_up.F) {          // my top H plane -> the upper slab's lower halo (its local plane 0), then its flag_h
        lk.up.F = (T *)s->peer_up.F; lk.up.fcs = s->peer_up.fcs; lk.up.set_stride = s->peer_up.set_stride; lk.up.kl = 0;
        lk.up.flag = (unsigned long long *)s->peer_up.sync + 1;
    }
    if (s->peer_down.F) {        // my bottom E plane -> the lower slab's upper halo (its last local plane), then its flag_e
        lk.down.F = (T *)s->peer_down.F; lk.down.fcs = s->peer_down.fcs; lk.down.set_stride = s->peer_down.set_stride;
        lk.down.kl = s->peer_down.nzl - 1; lk.down.flag = (unsigned long long *)s->peer_down.sync;
    }
    lk.flag_e = s->sync_dev; lk.flag_h = s->sync_dev + 1; lk.done = (unsigned int *)(s->sync_dev + 2); lk.err = s->flags + 1;
    lk.n_bnd[0] = s->tma.n_bnd[0] * s->g.n_sets * (NT / 32); lk.n_bnd[1] = s->tma.n_bnd[1] * s->g.n_sets * (NT / 32);
    // one block per SM: the whole shared memory as one ring, 255 registers (a two-blocks-per-SM build with 128 registers
    // spilled in the heavy E bodies and was slower; its instantiations doubled the compile time and were dropped)
    static const int cap_kb = tma_env("SJ_TMA_RING_KB", 208);
    const int cap = cap_kb * 1024;
    const size_t smem = (size_t)cap + SJ_RING_HEADER + 128;     // header, ring, alignment slack
    s->fan_main = st;
    auto kern = step_tma<T, NT, NB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TR(s, which == 0 ? "step_tma(H)" : which == 1 ? "step_tma(E)" : "step_tma", st, kern<<<l.grid, NT + 32, smem, st__>>>(p, bs, plan, lk, cap, s->kz0, s->kz1));
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}
