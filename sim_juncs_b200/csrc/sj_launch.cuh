// sj_launch.cuh -- kernel parameter blocks and the launch logic of one half-pass (work lists -> kernels, fan-out over
// side streams).  The heavy kernel templates are instantiated only by sj_launch_f64.cu / sj_launch_f32.cu, one
// translation unit per precision, so the library builds in parallel; sj_engine.cu includes this header for
// fill_params and the small kernels only.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "sj_kernels.cuh"

template <typename T>
static void fill_params(const sj_sim *s, KParams<T> &p) {
    memset(&p, 0, sizeof p);
    for (int d = 0; d < 3; ++d) p.n[d] = s->g.n[d];
    p.pitch = s->pitch; p.rows = s->rows; p.plane = s->plane; p.set_stride = s->set_stride;
    p.kz0 = s->kz0; p.nzl = s->nzl; p.n_sets = s->g.n_sets;
    p.F = (T *)s->F; p.fcs = s->set_stride * s->g.n_sets;
    for (int c = 0; c < 3; ++c) { p.E[c] = (T *)s->E[c]; p.H[c] = (T *)s->H[c]; p.mat[c] = s->mat[c]; p.sig[c] = (const T *)s->sigd[c]; p.siginv[c] = (const T *)s->siginvd[c]; }
    p.Pall = (T *)s->Pall; p.p_k0 = s->p_k0; p.p_nzp = s->p_nzp; p.p_set_stride = s->plane * s->p_nzp;
    p.p_comp_stride = p.p_set_stride * s->g.n_sets; p.n_slots = std::max(s->n_slots, 1);
    for (int q = 0; q < SJ_MAX_POLES; ++q) p.np_thr[q] = s->np_thr[q];
    p.mt_eps = (const T *)s->mt_eps; p.mt_chi = (const T *)s->mt_chi; p.mt_np = s->mt_np; p.mt_coef = (const T *)s->mt_coef; p.first_disp = s->first_disp;
    p.courant = (T)s->g.courant;
    p.n_src = (int)s->srcs.size();
    for (int q = 0; q < p.n_src; ++q) { p.src_klo[q] = s->srcs[q].lo[2]; p.src_khi[q] = s->srcs[q].hi[2]; }
    p.srcd = (const SrcDev<T> *)s->src_dev;
    p.drive = (const T *)s->drive;
    p.step = s->step_dev;
}

template <typename T>
static void fill_box(const sj_sim::Box &B, PmlBox<T> &b, int n_sets) {
    b.bcs = B.bset * n_sets;
    for (int d = 0; d < 3; ++d) { b.lo[d] = B.lo[d]; b.hi[d] = B.hi[d]; }
    b.bx = B.bx; b.by = B.by; b.bpitch = B.bpitch; b.bplane = B.bplane; b.bset = B.bset;
    for (int c = 0; c < 3; ++c) { b.D[c] = (T *)B.D[c]; b.B[c] = (T *)B.B[c]; b.UD[c] = (T *)B.UD[c]; b.UB[c] = (T *)B.UB[c]; }
}

// debug timeline: TR(s, name, stream, launch-expression)
#define TR(s, nm, strm, ...)                                                                        \
    do {                                                                                           \
        cudaStream_t st__ = (strm);                                                                \
        if ((s)->trace_on) { cudaEvent_t a__; cudaEventCreate(&a__); cudaEventRecord(a__, st__); (s)->tr_ev.push_back(a__); (s)->tr_name.push_back(nm); } \
        __VA_ARGS__;                                                                               \
        if ((s)->trace_on) { cudaEvent_t b__; cudaEventCreate(&b__); cudaEventRecord(b__, st__); (s)->tr_ev.push_back(b__); } \
    } while (0)

// round-robin over the main stream and the side streams between fan_begin / fan_end
static cudaStream_t fan_stream(sj_sim *s, int cls = 1) {
    // cls 0: the big interior kernels get the main stream / first side stream to themselves;
    // cls 1: PML lists round-robin over the remaining side streams
    if (!s->fan_on) return s->fan_main;
    static const bool dedicated = getenv("SJ_FAN_SHARED") == NULL;
    if (!dedicated || s->n_aux < 3) {
        const int q = s->fan_next++ % (s->n_aux + 1);
        return q == 0 ? s->fan_main : s->aux[q - 1];
    }
    if (cls == 0) { const int q = s->fan_int++ % 2; return q == 0 ? s->fan_main : s->aux[0]; }
    const int q = s->fan_next++ % (s->n_aux - 1);
    return s->aux[1 + q];
}
static void fan_begin(sj_sim *s, cudaStream_t st) {
    s->fan_main = st; s->fan_next = 0; s->fan_int = 0;
    if (!s->fan_on) return;
    cudaEventRecord(s->ev_fork, st);
    for (int a = 0; a < s->n_aux; ++a) cudaStreamWaitEvent(s->aux[a], s->ev_fork, 0);
}
static void fan_end(sj_sim *s) {
    if (!s->fan_on) return;
    for (int a = 0; a < s->n_aux; ++a) { cudaEventRecord(s->ev_join[a], s->aux[a]); cudaStreamWaitEvent(s->fan_main, s->ev_join[a], 0); }
}

template <typename T, int V, int LX>
static void launch_interior(sj_sim *s, const KParams<T> &p, int which, int k_begin, int k_end, cudaStream_t st) {
    IntGeom g; dim3 grd; sj_interior_geom(s, k_begin, k_end, g, grd);
    if (g.nzc <= 0 || !grd.x || !grd.y) return;
    if (which == 0) {
        static const bool pipe = getenv("SJ_H_PIPE") ? atoi(getenv("SJ_H_PIPE")) != 0 : true;
        if (pipe) TR(s, "h_interior_pipe<LX>", fan_stream(s, 0), h_interior_pipe<T, V, LX><<<grd, 256, 0, st__>>>(p, g, k_begin, k_end));
        else TR(s, "h_interior<LX>", fan_stream(s, 0), h_interior<T, V, LX><<<grd, 256, 0, st__>>>(p, g, k_begin, k_end));
        s->launches++; return;
    }
    if (s->il_int[0].n) { TR(s, "e_interior<LX, 0>", fan_stream(s, 0), e_interior<T, V, LX, 0><<<s->il_int[0].n, 256, 0, st__>>>(p, g, s->il_int[0].dev, k_begin, k_end)); s->launches++; }
    static const bool stg = getenv("SJ_NO_STAGE") == NULL;
    for (int cls = 1; cls < 4; ++cls) {       // 1: mixed tiles; 2 / 3: one material with exactly 1 / 2 poles
        if (!s->il_int[cls].n) continue;
        const int n = s->il_int[cls].n; const WorkItem *d = s->il_int[cls].dev;
        if (stg && s->n_slots <= 2) {
            // cp.async-staged variants: 2 stages x (8 + 6 NS) slots x 256 threads x 16 B of dynamic shared memory
            const int ns = cls == 1 ? std::max(s->n_slots, 1) : cls - 1;
            const size_t smem = (size_t)2 * (8 + 6 * ns) * 256 * 16;
            auto k1 = e_interior_stg<T, V, LX, 1, false>; auto k2 = e_interior_stg<T, V, LX, 2, false>;
            auto u1 = e_interior_stg<T, V, LX, 1, true>; auto u2 = e_interior_stg<T, V, LX, 2, true>;
            auto kf = cls == 1 ? (ns == 1 ? k1 : k2) : (ns == 1 ? u1 : u2);
            cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            TR(s, cls == 1 ? "e_interior_stg" : (ns == 1 ? "e_interior_stg<uni 1>" : "e_interior_stg<uni 2>"), fan_stream(s, 0),
               kf<<<n, 256, smem, st__>>>(p, g, d, k_begin, k_end));
        }
        else if (s->n_slots <= 1) TR(s, "e_interior<LX, 1>", fan_stream(s, 0), e_interior<T, V, LX, 1><<<n, 256, 0, st__>>>(p, g, d, k_begin, k_end));
        else if (s->n_slots == 2) TR(s, "e_interior<LX, 2>", fan_stream(s, 0), e_interior<T, V, LX, 2><<<n, 256, 0, st__>>>(p, g, d, k_begin, k_end));
        else TR(s, "e_interior<LX, 4>", fan_stream(s, 0), e_interior<T, V, LX, 4><<<n, 256, 0, st__>>>(p, g, d, k_begin, k_end));
        s->launches++;
    }
}

template <typename T, int V, int LX, bool FACE>
static void launch_e_pml(sj_sim *s, const KParams<T> &p, const PmlBoxSet<T> &bs, const ItemList (&L)[4], int k_begin, int k_end,
                         cudaStream_t st) {
    if (L[0].n) { TR(s, "e_pml_tile<LX, 0, FACE>", fan_stream(s), e_pml_tile<T, V, LX, 0, FACE, true><<<L[0].n, 256, 0, st__>>>(p, bs, L[0].dev, k_begin, k_end)); s->launches++; }
    if (L[1].n) {
        if (s->n_slots <= 1) TR(s, "e_pml_tile<LX, 1, FACE>", fan_stream(s), e_pml_tile<T, V, LX, 1, FACE, false><<<L[1].n, 256, 0, st__>>>(p, bs, L[1].dev, k_begin, k_end));
        else if (s->n_slots == 2) TR(s, "e_pml_tile<LX, 2, FACE>", fan_stream(s), e_pml_tile<T, V, LX, 2, FACE, false><<<L[1].n, 256, 0, st__>>>(p, bs, L[1].dev, k_begin, k_end));
        else TR(s, "e_pml_tile<LX, 4, FACE>", fan_stream(s), e_pml_tile<T, V, LX, 4, FACE, false><<<L[1].n, 256, 0, st__>>>(p, bs, L[1].dev, k_begin, k_end));
        s->launches++;
    }
    // one material with exactly 1 / 2 poles: coefficients in registers, no material bytes
    if (L[2].n) { TR(s, "e_pml_tile<LX, 1, FACE, uni>", fan_stream(s), e_pml_tile<T, V, LX, 1, FACE, true><<<L[2].n, 256, 0, st__>>>(p, bs, L[2].dev, k_begin, k_end)); s->launches++; }
    if (L[3].n) { TR(s, "e_pml_tile<LX, 2, FACE, uni>", fan_stream(s), e_pml_tile<T, V, LX, 2, FACE, true><<<L[3].n, 256, 0, st__>>>(p, bs, L[3].dev, k_begin, k_end)); s->launches++; }
}

template <typename T, int V, int LX>
static void launch_pml_lx(sj_sim *s, const KParams<T> &p, const PmlBoxSet<T> &bs, int which, int wn, int k_begin, int k_end,
                          cudaStream_t st) {
    if (which == 0) {
        if (s->il_h[0][wn].n) { TR(s, "h_pml_tile<LX, false>", fan_stream(s), h_pml_tile<T, V, LX, false><<<s->il_h[0][wn].n, 256, 0, st__>>>(p, bs, s->il_h[0][wn].dev, k_begin, k_end)); s->launches++; }
        if (s->il_h[1][wn].n) { TR(s, "h_pml_tile<LX, true>", fan_stream(s), h_pml_tile<T, V, LX, true><<<s->il_h[1][wn].n, 256, 0, st__>>>(p, bs, s->il_h[1][wn].dev, k_begin, k_end)); s->launches++; }
        return;
    }
    launch_e_pml<T, V, LX, false>(s, p, bs, s->il_pml[0][wn], k_begin, k_end, st);
    launch_e_pml<T, V, LX, true>(s, p, bs, s->il_pml[1][wn], k_begin, k_end, st);
}

template <typename T, int V>
static void launch_pml(sj_sim *s, const KParams<T> &p, int which, int k_begin, int k_end, cudaStream_t st) {
    PmlBoxSet<T> bs;
    memset(&bs, 0, sizeof bs);
    for (size_t bi = 0; bi < s->boxes.size(); ++bi) fill_box(s->boxes[bi], bs.b[bi], s->g.n_sets);
    if (s->pml_v != V) {       // half-width vectors: wide tiles 32 lanes, narrow tiles 16
        launch_pml_lx<T, V / 2, 32>(s, p, bs, which, 0, k_begin, k_end, st);
        launch_pml_lx<T, V / 2, 16>(s, p, bs, which, 1, k_begin, k_end, st);
        return;
    }
    if (s->pml_lx == 32) launch_pml_lx<T, V, 32>(s, p, bs, which, 0, k_begin, k_end, st);
    else launch_pml_lx<T, V, 16>(s, p, bs, which, 0, k_begin, k_end, st);
    launch_pml_lx<T, V, 8>(s, p, bs, which, 1, k_begin, k_end, st);
}

template <typename T, int V>
static int launch_pass(sj_sim *s, int which, int k_begin, int k_end, cudaStream_t st) {
    KParams<T> p; fill_params(s, p);
    k_begin = std::max(k_begin, s->kz0); k_end = std::min(k_end, s->kz1);
    if (k_begin >= k_end) return 0;
    fan_begin(s, st);
    static const bool pml_first = getenv("SJ_PML_LAST") == NULL;   // small latency-bound PML kernels first, interior fills in
    if (pml_first) launch_pml<T, V>(s, p, which, k_begin, k_end, st);
    if (s->int_lx == 32) launch_interior<T, V, 32>(s, p, which, k_begin, k_end, st);
    else if (s->int_lx == 16) launch_interior<T, V, 16>(s, p, which, k_begin, k_end, st);
    else launch_interior<T, V, 8>(s, p, which, k_begin, k_end, st);
    if (!pml_first) launch_pml<T, V>(s, p, which, k_begin, k_end, st);
    fan_end(s);
    CK(cudaGetLastError());
    return 0;
}


template <typename T, int V>
static int profile_impl(sj_sim *s, int reps, double out[4]) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    KParams<T> p; fill_params(s, p);
    const bool fan_saved = s->fan_on; s->fan_main = s->stream;
    for (int fam = 0; fam < 4; ++fam) {
        // families of several kernels (the E interior lists, the PML lists) are timed as they run in a step: fanned
        // over their streams
        s->fan_on = fan_saved && fam >= 1;
        out[fam] = 0.0;
        for (int rep = -2; rep < reps; ++rep) {          // two untimed warm-up launches
            if (rep == 0) CK(cudaEventRecord(e0, s->stream));
            if (fam < 2) {
                fan_begin(s, s->stream);
                if (s->int_lx == 32) launch_interior<T, V, 32>(s, p, fam, s->kz0, s->kz1, s->stream);
                else if (s->int_lx == 16) launch_interior<T, V, 16>(s, p, fam, s->kz0, s->kz1, s->stream);
                else launch_interior<T, V, 8>(s, p, fam, s->kz0, s->kz1, s->stream);
                fan_end(s);
            } else {
                fan_begin(s, s->stream);
                launch_pml<T, V>(s, p, fam - 2, s->kz0, s->kz1, s->stream);
                fan_end(s);
            }
        }
        CK(cudaEventRecord(e1, s->stream));
        CK(cudaEventSynchronize(e1));
        float f = 0; CK(cudaEventElapsedTime(&f, e0, e1));
        out[fam] = f / std::max(reps, 1);
    }
    s->fan_on = fan_saved;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CK(cudaGetLastError());
    return SJ_OK;
}

