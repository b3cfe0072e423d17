// sj_tma_launch.cuh -- launch logic of the TMA-staged kernels for one precision (instantiated by sj_tma_f64.cu / _f32.cu)
#pragma once
#include "sj_launch.cuh"
#include "sj_tma.cuh"

template <typename T>
static void tma_plan(const sj_sim *s, const TmaList &l, TmaPlan &plan) {
    plan.maps = (const CUtensorMap *)s->tma.maps;
    for (int i = 0; i < SJ_TMA_MAX_SHAPES; ++i) plan.shape[i] = s->tma.shapes[i];
    plan.items = l.items; plan.blk_first = l.first;
}

#define SJ_TMA_LAUNCH(s, nm, cls, kern, stage_bytes, nst, nt, list, p, bs)                                          \
    do {                                                                                                           \
        const TmaList &l__ = (list);                                                                               \
        if (l__.n_items) {                                                                                         \
            TmaPlan plan__; tma_plan<T>(s, l__, plan__);                                                           \
            const size_t smem__ = (size_t)(stage_bytes) * (nst) + 256;                                             \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem__);                  \
            TR(s, nm, fan_stream(s, cls), kern<<<l__.grid, (nt) + 32, smem__, st__>>>(p, bs, plan__, s->kz0, s->kz1)); \
            s->launches++;                                                                                         \
        }                                                                                                          \
    } while (0)

static int tma_env(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

template <typename T>
static int tma_pass(sj_sim *s, int which, cudaStream_t st) {
    KParams<T> p; fill_params(s, p);
    PmlBoxSet<T> bs; memset(&bs, 0, sizeof bs);
    for (size_t bi = 0; bi < s->boxes.size(); ++bi) fill_box(s->boxes[bi], bs.b[bi], s->g.n_sets);
    constexpr int NT = 224, NG = 96;        // consumer threads: interior / face tiles, edge / corner tiles
    typedef Slots<NT> SL;
    typedef Slots<NG> SLG;
    const TmaState &t = s->tma;
    fan_begin(s, st);
    if (which == 0) {
        static const int nst_a = tma_env("SJ_TMA_HNST", 4), nst_g = tma_env("SJ_TMA_HGNST", 4);
        constexpr int SA = 3 * SL::HALO + 4 * SL::OWN, SG = 3 * SLG::HALO + 9 * SLG::OWN;
        if (nst_a >= 6) SJ_TMA_LAUNCH(s, "h_tma<A>", 0, (h_tma<T, NT, 6, false, 1>), SA, 6, NT, t.h[0], p, bs);
        else if (nst_a == 4) SJ_TMA_LAUNCH(s, "h_tma<A>", 0, (h_tma<T, NT, 4, false, 2>), SA, 4, NT, t.h[0], p, bs);
        else SJ_TMA_LAUNCH(s, "h_tma<A>", 0, (h_tma<T, NT, 3, false, 2>), SA, 3, NT, t.h[0], p, bs);
        if (nst_g >= 4) SJ_TMA_LAUNCH(s, "h_tma<general>", 1, (h_tma<T, NG, 4, true, 2>), SG, 4, NG, t.h[1], p, bs);
        else SJ_TMA_LAUNCH(s, "h_tma<general>", 1, (h_tma<T, NG, 2, true, 2>), SG, 2, NG, t.h[1], p, bs);
    } else {
        static const int nst_0 = tma_env("SJ_TMA_E0NST", 4), nst_1 = tma_env("SJ_TMA_E2NST", 2);
        typedef EStage<NT, 0, false> A0; typedef EStage<NT, 1, false> A1; typedef EStage<NT, 2, false> A2;
        typedef EStage<NG, 0, true> G0; typedef EStage<NG, 1, true> G1; typedef EStage<NG, 2, true> G2;
        // interior + face tiles, heaviest class first
        if (nst_1 >= 4) SJ_TMA_LAUNCH(s, "e_tma<A, uni 1>", 0, (e_tma<T, NT, 4, 1, true, false, 1>), A1::BYTES, 4, NT, t.e[0][2], p, bs);
        else if (nst_1 == 3) SJ_TMA_LAUNCH(s, "e_tma<A, uni 1>", 0, (e_tma<T, NT, 3, 1, true, false, 1>), A1::BYTES, 3, NT, t.e[0][2], p, bs);
        else SJ_TMA_LAUNCH(s, "e_tma<A, uni 1>", 0, (e_tma<T, NT, 2, 1, true, false, 2>), A1::BYTES, 2, NT, t.e[0][2], p, bs);
        if (nst_0 >= 6) SJ_TMA_LAUNCH(s, "e_tma<A, 0>", 0, (e_tma<T, NT, 6, 0, true, false, 1>), A0::BYTES, 6, NT, t.e[0][0], p, bs);
        else if (nst_0 == 4) SJ_TMA_LAUNCH(s, "e_tma<A, 0>", 0, (e_tma<T, NT, 4, 0, true, false, 2>), A0::BYTES, 4, NT, t.e[0][0], p, bs);
        else SJ_TMA_LAUNCH(s, "e_tma<A, 0>", 0, (e_tma<T, NT, 3, 0, true, false, 2>), A0::BYTES, 3, NT, t.e[0][0], p, bs);
        if (s->n_slots <= 1) SJ_TMA_LAUNCH(s, "e_tma<A, mixed 1>", 1, (e_tma<T, NT, 2, 1, false, false, 1>), A1::BYTES, 2, NT, t.e[0][1], p, bs);
        else SJ_TMA_LAUNCH(s, "e_tma<A, mixed 2>", 1, (e_tma<T, NT, 2, 2, false, false, 1>), A2::BYTES, 2, NT, t.e[0][1], p, bs);
        SJ_TMA_LAUNCH(s, "e_tma<A, uni 2>", 1, (e_tma<T, NT, 2, 2, true, false, 1>), A2::BYTES, 2, NT, t.e[0][3], p, bs);
        // edge / corner tiles, by material class
        SJ_TMA_LAUNCH(s, "e_tma<general, 0>", 1, (e_tma<T, NG, 4, 0, true, true, 2>), G0::BYTES, 4, NG, t.e[1][0], p, bs);
        SJ_TMA_LAUNCH(s, "e_tma<general, uni 1>", 1, (e_tma<T, NG, 3, 1, true, true, 2>), G1::BYTES, 3, NG, t.e[1][2], p, bs);
        if (s->n_slots <= 1) SJ_TMA_LAUNCH(s, "e_tma<general, mixed 1>", 1, (e_tma<T, NG, 3, 1, false, true, 2>), G1::BYTES, 3, NG, t.e[1][1], p, bs);
        else SJ_TMA_LAUNCH(s, "e_tma<general, mixed 2>", 1, (e_tma<T, NG, 3, 2, false, true, 1>), G2::BYTES, 3, NG, t.e[1][1], p, bs);
        SJ_TMA_LAUNCH(s, "e_tma<general, uni 2>", 1, (e_tma<T, NG, 3, 2, true, true, 1>), G2::BYTES, 3, NG, t.e[1][3], p, bs);
    }
    fan_end(s);
    CK(cudaGetLastError());
    return 0;
}
