// sj_tma_launch.cuh -- launch logic of the TMA-staged kernels for one precision (instantiated by sj_tma_f64.cu / _f32.cu)
#pragma once
#include "sj_launch.cuh"
#include "sj_tma.cuh"

template <typename T>
static void tma_plan(const sj_sim *s, const TmaList &l, TmaPlan &plan) {
    plan.maps = (const CUtensorMap *)s->tma.maps;
    for (int i = 0; i < SJ_TMA_MAX_SHAPES; ++i) plan.shape[i] = s->tma.shapes[i];
    plan.items = l.items; plan.n_items = l.n_items; plan.queue = l.first; plan.tick = nullptr;
}

static int tma_env(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

// one persistent kernel per half-pass: one block per SM, the whole shared memory as its staging ring
template <typename T>
static int tma_pass(sj_sim *s, int which, cudaStream_t st, bool tick) {
    KParams<T> p; fill_params(s, p);
    PmlBoxSet<T> bs; memset(&bs, 0, sizeof bs);
    for (size_t bi = 0; bi < s->boxes.size(); ++bi) fill_box(s->boxes[bi], bs.b[bi], s->g.n_sets);
    constexpr int NT = 224, NB = 12;
    const TmaList &l = which == 0 ? s->tma.h[0] : s->tma.e[0][0];
    if (!l.n_items) return 0;
    TmaPlan plan; tma_plan<T>(s, l, plan);
    if (which == 1 && tick) plan.tick = s->step_dev;
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
    // SJ_TMA_BLK = blocks per SM (1: the whole shared memory as one ring, 255 registers; 2: two rings, 128 registers)
    static const int blk = tma_env("SJ_TMA_BLK", 1);
    static const int cap_kb = tma_env("SJ_TMA_RING_KB", blk == 2 ? 104 : 208);
    const int cap = cap_kb * 1024;
    const size_t smem = (size_t)cap + 768;
    s->fan_main = st;
    if (which == 0) {
        auto kern = blk == 2 ? h_tma<T, NT, NB, 2> : h_tma<T, NT, NB, 1>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        TR(s, "h_tma", st, kern<<<l.grid, NT + 32, smem, st__>>>(p, bs, plan, lk, cap, s->kz0, s->kz1));
    } else {
        auto kern = blk == 2 ? e_tma<T, NT, NB, 2> : e_tma<T, NT, NB, 1>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        TR(s, "e_tma", st, kern<<<l.grid, NT + 32, smem, st__>>>(p, bs, plan, lk, cap, s->kz0, s->kz1));
    }
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}
