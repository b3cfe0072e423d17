// sj_engine.cu -- host runtime of the sim_juncs_b200 engine: grid/PML/source/monitor set-up
// (what meep::structure / meep::fields do at construction for the reference, SURVEY.md 8a rows
// A7, A8, M4, M8), device memory, and the step loop of bound_geom::run (src/disp.cpp:719-741).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <complex>

#include "sj_launch.cuh"

static std::string g_create_err;

static void graph_drop(sj_sim *s) {
    if (s->step_graph) { cudaGraphExecDestroy(s->step_graph); s->step_graph = NULL; }
}

static int fail(sj_sim *s, int code, const char *msg) {
    if (s) s->err = msg; else g_create_err = msg;
    return code;
}

extern "C" int sj_version(void) { return 100; }
extern "C" const char *sj_last_error(const sj_sim *s) { return s ? s->err.c_str() : g_create_err.c_str(); }

// meep structure_chunk::use_pml: quadratic profile, sigma sampled per half pixel, sig = sigma dt/2
static void build_pml_table(std::vector<double> &sig, int n, double a, double dt, double thick, double R) {
    sig.assign(2 * n + 2, 0.0);
    if (thick <= 0) return;
    const double inva = 1.0 / a;
    const double prefac = -log(R) / (4 * thick * (1.0 / 3.0));
    const double bloc[2] = {0.0, (2 * n) * (0.5 * inva)};
    for (int side = 0; side < 2; ++side)
        for (int i = 0; i < 2 * n + 2; ++i) {
            const double here = i * 0.5 * inva;
            const double x = thick - (0.5 * inva) * ((int)(fabs(bloc[side] - here) * a * 2 + 0.5));
            if (x > 0) {
                const double u = x / thick;
                sig[i] = 0.5 * dt * (prefac * u * u);
            }
        }
}

template <typename T>
static int upload_vec(sj_sim *s, const std::vector<double> &v, void **dev) {
    std::vector<T> tmp(v.size());
    for (size_t i = 0; i < v.size(); ++i) tmp[i] = (T)v[i];
    CK(cudaMalloc(dev, std::max<size_t>(tmp.size(), 1) * sizeof(T)));
    if (!tmp.empty()) CK(cudaMemcpy(*dev, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
static int upload_sig(sj_sim *s, int d) {
    // the device copies carry 16 entries of padding (sigma 0): the last vector of a row reads the table at the padding
    // columns i > n of the grid as well (found by compute-sanitizer memcheck)
    std::vector<double> sg(s->sig[d]), inv(s->sig[d].size());
    for (size_t i = 0; i < inv.size(); ++i) inv[i] = 1 / (1.0 + s->sig[d][i]);   // meep: siginv = 1/(kap+sig)
    sg.resize(sg.size() + 16, 0.0); inv.resize(inv.size() + 16, 1.0);
    int rc = s->prec == SJ_F64 ? upload_vec<double>(s, inv, &s->siginvd[d]) : upload_vec<float>(s, inv, &s->siginvd[d]);
    if (rc) return rc;
    return s->prec == SJ_F64 ? upload_vec<double>(s, sg, &s->sigd[d]) : upload_vec<float>(s, sg, &s->sigd[d]);
}

static int alloc_zero(sj_sim *s, void **p, size_t bytes) {
    CK(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    CK(cudaMemset(*p, 0, std::max<size_t>(bytes, 16)));
    CK(cudaDeviceSynchronize());        // the legacy-stream fill is not ordered with the simulation's non-blocking stream
    return 0;
}

extern "C" int sj_create(const sj_grid *g, sj_sim **out) {
    if (!g || !out) return fail(NULL, SJ_ERR_ARG, "null argument");
    if (g->n[0] < 2 || g->n[1] < 2 || g->n[2] < 2 || g->a <= 0) return fail(NULL, SJ_ERR_ARG, "bad grid");
    if (g->n_sets < 1 || g->n_sets > 64) return fail(NULL, SJ_ERR_ARG, "n_sets out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(NULL, SJ_ERR_CUDA, "no CUDA device: libsimjuncs_b200 has no CPU fallback");
    sj_sim *s = new sj_sim();
    s->g = *g;
    if (g->device >= 0 && cudaSetDevice(g->device) != cudaSuccess) { delete s; return fail(NULL, SJ_ERR_CUDA, "cudaSetDevice failed"); }
    if (g->device < 0) cudaGetDevice(&s->g.device);
    s->prec = g->precision == SJ_F32 ? SJ_F32 : SJ_F64;
    s->esz = s->prec == SJ_F64 ? 8 : 4;
    s->g.courant = g->courant > 0 ? g->courant : 0.5;
    s->g.pml_R = g->pml_R > 0 ? g->pml_R : 1e-15;
    s->inva = 1.0 / g->a;
    s->dt = s->g.courant * s->inva;
    s->kz0 = g->kz0; s->kz1 = g->kz1;
    if (s->kz0 == 0 && s->kz1 == 0) s->kz1 = g->n[2] + 1;
    if (s->kz0 < 0 || s->kz1 > g->n[2] + 1 || s->kz0 >= s->kz1) { delete s; return fail(NULL, SJ_ERR_ARG, "bad z-slab"); }
    s->pitch = ((g->n[0] + 1 + 31) / 32) * 32;
    s->rows = g->n[1] + 1;
    s->nzl = s->kz1 - s->kz0 + 2;
    s->plane = (long long)s->pitch * s->rows;
    s->set_stride = s->plane * s->nzl;
    s->step_graph = NULL; s->graph_launches = 0;
    s->materials_set = false; s->n_slots = 0; s->drive = NULL; s->drive_steps = 0; s->drive_dirty = true;
    s->n_mon = 0; s->mon_idx = NULL; s->mon_w = NULL; s->series = NULL; s->series_cap = 0; s->n_samples = 0;
    s->steps_done = 0; s->launches = 0; s->pole_points = 0; s->pml_cells = 0;
    s->mt_chi = s->mt_coef = NULL; s->mt_np = NULL;
    s->F = NULL;
    for (int c = 0; c < 3; ++c) { s->E[c] = s->H[c] = NULL; s->mat[c] = s->masks[c] = NULL; s->sigd[c] = s->siginvd[c] = NULL; }
    s->flags_int = NULL; s->mt_eps = NULL; s->src_dev = NULL; s->pml_lx = 32; s->pml_lx_n = 8; s->pml_v = 2;
    for (int a = 0; a < 4; ++a) { s->il_int[a].dev = NULL; s->il_int[a].n = 0; }
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { s->il_h[a][b].dev = NULL; s->il_h[a][b].n = 0; for (int c = 0; c < 4; ++c) { s->il_pml[a][b][c].dev = NULL; s->il_pml[a][b][c].n = 0; } }
    s->first_disp = 0; s->int_lx = 32; s->int_zchunk = 16;
    for (int i = 0; i < 256; ++i) s->lut_inv[i] = (uint8_t)i;
    s->Pall = NULL;
    for (int q = 0; q < SJ_MAX_POLES; ++q) s->np_thr[q] = 1 << 30;
    for (int q = 0; q < SJ_MAX_SRC; ++q) for (int c = 0; c < 3; ++c) s->srcw[q][c] = NULL;
    *out = s;

    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s->ev_a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->ev_b, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    {   // side streams; the ones that carry the PML lists (aux[1..]) can be given scheduling priority over the interior kernels
        int pr_lo = 0, pr_hi = 0;
        cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi);
        const bool prio = getenv("SJ_PML_PRIO") ? atoi(getenv("SJ_PML_PRIO")) != 0 : false;
        for (int a = 0; a < SJ_N_AUX; ++a) {
            CK(cudaStreamCreateWithPriority(&s->aux[a], cudaStreamNonBlocking, (prio && a >= 1) ? pr_hi : pr_lo));
            CK(cudaEventCreateWithFlags(&s->ev_join[a], cudaEventDisableTiming));
        }
    }
    s->fan_on = getenv("SJ_NO_FAN") == NULL; s->fan_next = 0; s->fan_main = s->stream;
    s->trace_on = false;
    s->n_aux = getenv("SJ_N_AUX") ? std::min(std::max(atoi(getenv("SJ_N_AUX")), 1), SJ_N_AUX) : 5;
    const int n[3] = {g->n[0], g->n[1], g->n[2]};
    for (int d = 0; d < 3; ++d) {
        build_pml_table(s->sig[d], n[d], g->a, s->dt, g->pml_thickness, s->g.pml_R);
        int rc = upload_sig(s, d); if (rc) return rc;
        // interior = index range where both half-pixel samples have sigma == 0 and every
        // component is updatable with all neighbours in range: [max(.,1), min(., n))
        int lo = 1, hi = n[d];
        while (lo < n[d] && (s->sig[d][2 * lo] != 0.0 || s->sig[d][2 * lo + 1] != 0.0)) ++lo;
        while (hi > lo && (s->sig[d][2 * (hi - 1)] != 0.0 || s->sig[d][2 * (hi - 1) + 1] != 0.0)) --hi;
        if (d == 0) { lo = ((lo + 3) / 4) * 4; hi = (hi / 4) * 4; if (hi < lo) hi = lo; }
        s->lo[d] = lo; s->hi[d] = hi;
    }
    {   // interior tiling: lanes along x per warp chosen to waste the fewest lanes
        const int V = s->prec == SJ_F64 ? 2 : 4, ni = std::max(s->hi[0] - s->lo[0], 1);
        double best = -1;
        for (int lx = 32; lx >= 8; lx /= 2) {
            const int tw = lx * V;
            const double eff = (double)ni / (double)(((ni + tw - 1) / tw) * tw);
            if (eff > best + 0.04) { best = eff; s->int_lx = lx; }
        }
        if (getenv("SJ_INT_LX")) s->int_lx = atoi(getenv("SJ_INT_LX"));
        // planes per block: about 16, evened out so the last chunk is not a stub (161 planes -> 11 x 15, not 10 x 16 + 1)
        const int nk = std::max(std::min(s->hi[2], s->kz1) - std::max(s->lo[2], s->kz0), 1);
        const int target = getenv("SJ_ZCHUNK") ? std::max(atoi(getenv("SJ_ZCHUNK")), 1) : 16;
        const int nzc = (nk + target - 1) / target;
        s->int_zchunk = (nk + nzc - 1) / nzc;
    }
    const size_t fbytes = (size_t)s->set_stride * g->n_sets * s->esz;
    {
        int rc = alloc_zero(s, &s->F, 6 * fbytes); if (rc) return rc;
        uint8_t *mall; rc = alloc_zero(s, (void **)&mall, 3 * (size_t)s->set_stride); if (rc) return rc;
        for (int c = 0; c < 3; ++c) {
            s->E[c] = (char *)s->F + (size_t)c * fbytes;
            s->H[c] = (char *)s->F + (size_t)(3 + c) * fbytes;
            s->mat[c] = mall + (size_t)c * s->set_stride;
        }
    }
    // PML shell boxes (global index boxes [lo,hi), k clipped to the owned slab)
    {
        const int L[3] = {s->lo[0], s->lo[1], s->lo[2]}, Hh[3] = {s->hi[0], s->hi[1], s->hi[2]};
        const int N1[3] = {n[0] + 1, n[1] + 1, n[2] + 1};
        int bl[6][3], bh[6][3];
        // z-low, z-high: full xy ; y-low, y-high: z interior ; x-low, x-high: y,z interior
        int q = 0;
        bl[q][0] = 0; bh[q][0] = N1[0]; bl[q][1] = 0; bh[q][1] = N1[1]; bl[q][2] = 0; bh[q][2] = L[2]; ++q;
        bl[q][0] = 0; bh[q][0] = N1[0]; bl[q][1] = 0; bh[q][1] = N1[1]; bl[q][2] = Hh[2]; bh[q][2] = N1[2]; ++q;
        bl[q][0] = 0; bh[q][0] = N1[0]; bl[q][1] = 0; bh[q][1] = L[1]; bl[q][2] = L[2]; bh[q][2] = Hh[2]; ++q;
        bl[q][0] = 0; bh[q][0] = N1[0]; bl[q][1] = Hh[1]; bh[q][1] = N1[1]; bl[q][2] = L[2]; bh[q][2] = Hh[2]; ++q;
        bl[q][0] = 0; bh[q][0] = L[0]; bl[q][1] = L[1]; bh[q][1] = Hh[1]; bl[q][2] = L[2]; bh[q][2] = Hh[2]; ++q;
        bl[q][0] = Hh[0]; bh[q][0] = N1[0]; bl[q][1] = L[1]; bh[q][1] = Hh[1]; bl[q][2] = L[2]; bh[q][2] = Hh[2]; ++q;
        // Every shell box is cut into rectangles: the part with a single non-zero sigma ("face", kind = its normal
        // direction) and the frame around it where sigmas overlap (kind 0: edges, corners).  Each rectangle is a region
        // with its own compact allocation: a face keeps only the normal component's D and B (2 arrays), a frame the full
        // [D(3) | B(3) | U_D(3) | U_B(3)].
        const int N1x = N1[0], N1y = N1[1];
        for (int b = 0; b < 6; ++b) {
            const int zlo = std::max(bl[b][2], s->kz0), zhi = std::min(bh[b][2], s->kz1);
            if (zlo >= zhi) continue;
            struct Rect { int i0, i1, j0, j1, kind; };
            std::vector<Rect> rects;
            const bool zbox = b < 2, ybox = b >= 2 && b < 4;
            if (zbox) {
                rects.push_back({L[0], Hh[0], L[1], Hh[1], 3});
                rects.push_back({0, N1x, 0, L[1], 0}); rects.push_back({0, N1x, Hh[1], N1y, 0});
                rects.push_back({0, L[0], L[1], Hh[1], 0}); rects.push_back({Hh[0], N1x, L[1], Hh[1], 0});
            } else if (ybox) {
                rects.push_back({L[0], Hh[0], bl[b][1], bh[b][1], 2});
                rects.push_back({0, L[0], bl[b][1], bh[b][1], 0}); rects.push_back({Hh[0], N1x, bl[b][1], bh[b][1], 0});
            } else rects.push_back({bl[b][0], bh[b][0], bl[b][1], bh[b][1], 1});
            for (const Rect &R : rects) {
                sj_sim::Box B;
                B.lo[0] = R.i0; B.hi[0] = R.i1; B.lo[1] = R.j0; B.hi[1] = R.j1; B.lo[2] = zlo; B.hi[2] = zhi;
                B.kind = R.kind; B.narr = R.kind ? 2 : 12;
                B.bx = B.hi[0] - B.lo[0]; B.by = B.hi[1] - B.lo[1]; B.bz = B.hi[2] - B.lo[2];
                if (B.bx <= 0 || B.by <= 0 || B.bz <= 0) continue;
                if ((int)s->boxes.size() >= SJ_N_PML_BOX) return fail(s, SJ_ERR_UNSUPPORTED, "too many PML regions");
                B.bpitch = ((B.bx + 3) / 4) * 4;
                B.bplane = (long long)B.bpitch * B.by;
                B.bset = B.bplane * B.bz;
                const size_t bytes = (size_t)B.bset * g->n_sets * s->esz;
                int rc = alloc_zero(s, &B.base, (size_t)B.narr * bytes); if (rc) return rc;
                for (int c = 0; c < 3; ++c) {
                    if (B.kind) {   // only component kind-1 exists: D at array 0, B at array 1; the other pointers are never used
                        const long long off = (long long)(c - (B.kind - 1)) * (long long)bytes;
                        B.D[c] = (char *)B.base + off; B.B[c] = (char *)B.base + (long long)bytes + off;
                        B.UD[c] = B.base; B.UB[c] = B.base;
                    } else {
                        B.D[c] = (char *)B.base + (size_t)c * bytes; B.B[c] = (char *)B.base + (size_t)(3 + c) * bytes;
                        B.UD[c] = (char *)B.base + (size_t)(6 + c) * bytes; B.UB[c] = (char *)B.base + (size_t)(9 + c) * bytes;
                    }
                }
                s->pml_cells += (double)B.bx * B.by * B.bz;
                s->pml_bytes += (double)B.narr * bytes;
                s->boxes.push_back(B);
            }
        }
    }
    // work lists of the tiled PML kernels: one entry per thread block.  Every box is cut into
    // rectangles: the part with a single non-zero sigma ("face", kind = normal direction) and the
    // frame around it where sigmas overlap (kind 0, general kernel).
    {
        // PML kernels are instruction-latency bound at one or two blocks per SM with full 128-bit vectors
        // (ncu: 2 warps per scheduler); half-width vectors halve the registers and double the resident warps
        const bool half = getenv("SJ_PML_HALF") ? atoi(getenv("SJ_PML_HALF")) != 0 : true;
        const int V = (s->prec == SJ_F64 ? 2 : 4) / (half ? 2 : 1);
        s->pml_v = V;
        s->pml_lx = half ? 32 : std::max(s->int_lx, 16);
        s->pml_lx_n = half ? 16 : 8;
        const int zc_face = getenv("SJ_ZC_FACE") ? atoi(getenv("SJ_ZC_FACE")) : 8, zc_gen = getenv("SJ_ZC_GEN") ? atoi(getenv("SJ_ZC_GEN")) : 4;
        const int L[3] = {s->lo[0], s->lo[1], s->lo[2]}, Hh[3] = {s->hi[0], s->hi[1], s->hi[2]};
        const int N1x = n[0] + 1, N1y = n[1] + 1;
        for (size_t bi = 0; bi < s->boxes.size(); ++bi) {
            const sj_sim::Box &B = s->boxes[bi];
            struct Rect { int i0, i1, j0, j1, kind; };
            std::vector<Rect> rects;
            rects.push_back({B.lo[0], B.hi[0], B.lo[1], B.hi[1], B.kind});
            for (const Rect &R : rects) {
                if (R.i1 <= R.i0 || R.j1 <= R.j0) continue;
                const bool nar = (R.i1 - R.i0 <= s->pml_lx_n * V);
                const int tw = (nar ? s->pml_lx_n : s->pml_lx) * V, th = (32 / (nar ? s->pml_lx_n : s->pml_lx)) * 8;
                const int zchunk = R.kind ? zc_face : zc_gen;   // short marches: these lists are small, keep every SM busy
                for (int q = 0; q < g->n_sets; ++q)
                    for (int kb = B.lo[2]; kb < B.hi[2]; kb += zchunk)
                        for (int j0 = R.j0; j0 < R.j1; j0 += th)
                            for (int i0 = R.i0; i0 < R.i1; i0 += tw) {
                                WorkItem w = {(int)bi, q, i0, j0, kb, std::min(kb + zchunk, B.hi[2]), 0, R.kind, R.i1, R.j1};
                                s->h_items[R.kind ? 1 : 0][nar ? 1 : 0].push_back(w);
                            }
            }
        }
        for (int f = 0; f < 2; ++f)
            for (int w = 0; w < 2; ++w) {
                const std::vector<WorkItem> &v = s->h_items[f][w];
                s->il_h[f][w].n = (int)v.size();
                CK(cudaMalloc((void **)&s->il_h[f][w].dev, std::max<size_t>(v.size(), 1) * sizeof(WorkItem)));
                if (!v.empty()) CK(cudaMemcpy(s->il_h[f][w].dev, v.data(), v.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
            }
    }
    {   // TMA-staged kernels (sj_tma.cuh): SJ_TMA bit 0 = H-pass, bit 1 = E-pass
        s->tma.mode = getenv("SJ_TMA") ? atoi(getenv("SJ_TMA")) : 3;
        int rc = sj_tma_build_geometry(s); if (rc) return rc;
    }
    // default material table: vacuum
    {
        sj_material vac; memset(&vac, 0, sizeof vac); vac.eps_inf = 1.0;
        int rc = sj_set_materials(s, 1, &vac, NULL, NULL, NULL); if (rc) return rc;
        s->materials_set = false;
    }
    int rc = alloc_zero(s, (void **)&s->step_dev, sizeof(long long)); if (rc) return rc;
    rc = alloc_zero(s, (void **)&s->flags, 4 * sizeof(int)); if (rc) return rc;
    rc = alloc_zero(s, (void **)&s->sync_dev, 4 * sizeof(unsigned long long)); if (rc) return rc;
    CK(cudaDeviceSynchronize());        // the zero fills above ran on the legacy stream; the simulation's stream is non-blocking
    return SJ_OK;
}

extern "C" void sj_destroy(sj_sim *s) {
    if (!s) return;
    cudaStreamSynchronize(s->stream);
    graph_drop(s);
    cudaFree(s->F); cudaFree(s->mat[0]);
    for (int c = 0; c < 3; ++c) { cudaFree(s->masks[c]); cudaFree(s->sigd[c]); cudaFree(s->siginvd[c]); }
    cudaFree(s->flags_int); cudaFree(s->mt_eps);
    for (int a = 0; a < 4; ++a) cudaFree(s->il_int[a].dev);
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { cudaFree(s->il_h[a][b].dev); for (int c = 0; c < 4; ++c) cudaFree(s->il_pml[a][b][c].dev); }
    cudaFree(s->Pall); cudaFree(s->src_dev);
    sj_tma_free(s);
    for (auto &B : s->boxes) cudaFree(B.base);
    for (int q = 0; q < SJ_MAX_SRC; ++q) for (int c = 0; c < 3; ++c) cudaFree(s->srcw[q][c]);
    cudaFree(s->mt_chi); cudaFree(s->mt_coef); cudaFree(s->mt_np); cudaFree(s->drive);
    cudaFree(s->mon_idx); cudaFree(s->mon_w); cudaFree(s->series); cudaFree(s->step_dev); cudaFree(s->flags);
    if (s->peer_up.ipc) { cudaIpcCloseMemHandle(s->peer_up.F); cudaIpcCloseMemHandle(s->peer_up.sync); }
    if (s->peer_down.ipc) { cudaIpcCloseMemHandle(s->peer_down.F); cudaIpcCloseMemHandle(s->peer_down.sync); }
    cudaFree(s->sync_dev);
    cudaEventDestroy(s->ev_a); cudaEventDestroy(s->ev_b); cudaEventDestroy(s->ev_fork);
    for (int a = 0; a < SJ_N_AUX; ++a) { cudaStreamDestroy(s->aux[a]); cudaEventDestroy(s->ev_join[a]); }
    cudaStreamDestroy(s->stream);
    delete s;
}

// ---- materials -----------------------------------------------------------------------------
// ADE coefficients folded in fp64 on the host (SURVEY hard part H4), following
// meep lorentzian_susceptibility::update_P:
//   P+ = g1inv * (P (2 - w0^2 dt^2 [lorentz]) - g1 P- + w0^2 dt^2 sigma E)
static int upload_material_table(sj_sim *s) {
    std::vector<double> chi(SJ_MAX_MAT, 1.0), coef((size_t)SJ_MAX_MAT * SJ_MAX_POLES * 3, 0.0);
    std::vector<int> np(SJ_MAX_MAT, 0);
    int slots = 0;
    for (size_t m = 0; m < s->mats_sorted.size(); ++m) {
        const sj_material &M = s->mats_sorted[m];
        chi[m] = 1.0 / M.eps_inf;
        np[m] = M.n_poles;
        if (s->present[m]) slots = std::max(slots, M.n_poles);   // only materials that occur in the slab
        for (int q = 0; q < M.n_poles; ++q) {
            const sj_pole &P = M.poles[q];
            const double omega2pi = 2 * M_PI * P.omega0, g2pi = P.gamma * 2 * M_PI;
            const double omega0dtsqr = omega2pi * omega2pi * s->dt * s->dt;
            const double gamma1inv = 1 / (1 + g2pi * s->dt / 2), gamma1 = (1 - g2pi * s->dt / 2);
            const double denom = P.drude ? 0 : omega0dtsqr;
            double *cf = &coef[(m * SJ_MAX_POLES + q) * 3];
            cf[0] = gamma1inv * (2 - denom);
            cf[1] = -gamma1inv * gamma1;
            cf[2] = gamma1inv * omega0dtsqr * P.sigma;
        }
    }
    cudaFree(s->mt_chi); cudaFree(s->mt_coef); cudaFree(s->mt_np); cudaFree(s->mt_eps);
    s->mt_chi = s->mt_coef = s->mt_eps = NULL; s->mt_np = NULL;
    std::vector<double> epsv(SJ_MAX_MAT, 1.0);
    for (size_t m = 0; m < s->mats_sorted.size(); ++m) epsv[m] = s->mats_sorted[m].eps_inf;
    { int rc0 = s->prec == SJ_F64 ? upload_vec<double>(s, epsv, &s->mt_eps) : upload_vec<float>(s, epsv, &s->mt_eps); if (rc0) return rc0; }
    int rc;
    if (s->prec == SJ_F64) { rc = upload_vec<double>(s, chi, &s->mt_chi); if (rc) return rc; rc = upload_vec<double>(s, coef, &s->mt_coef); }
    else { rc = upload_vec<float>(s, chi, &s->mt_chi); if (rc) return rc; rc = upload_vec<float>(s, coef, &s->mt_coef); }
    if (rc) return rc;
    CK(cudaMalloc((void **)&s->mt_np, SJ_MAX_MAT * sizeof(int)));
    CK(cudaMemcpy(s->mt_np, np.data(), SJ_MAX_MAT * sizeof(int), cudaMemcpyHostToDevice));
    // polarisation slots: [parity][slot][comp][set] arrays over the local planes [p_k0, p_k0 + p_nzp) that hold a pole
    // material (set by sj_finish_materials before this call); loads are skipped where the material has no pole
    if (slots != s->n_slots || !s->Pall || s->p_k0 != s->p_alloc_k0 || s->p_nzp != s->p_alloc_nzp) {
        cudaFree(s->Pall); s->Pall = NULL;
        const size_t bytes = (size_t)2 * std::max(slots, 1) * 3 * s->plane * s->p_nzp * s->g.n_sets * s->esz;
        rc = alloc_zero(s, &s->Pall, bytes); if (rc) return rc;
        s->p_alloc_k0 = s->p_k0; s->p_alloc_nzp = s->p_nzp; s->p_bytes = (double)bytes;
    }
    for (int q = 0; q < SJ_MAX_POLES; ++q) {
        s->np_thr[q] = 1 << 30;
        for (int m = (int)s->mats_sorted.size() - 1; m >= 0; --m) if (s->mats_sorted[m].n_poles > q) s->np_thr[q] = m;
    }
    s->n_slots = slots;
    return 0;
}

__global__ void count_pole_points(const uint8_t *mat, const int *np, long long plane, int pitch, int i0, int i1, int j0,
                                  int j1, int kl0, int kl1, unsigned long long *out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = plane * (kl1 - kl0);
    unsigned long long acc = 0;
    for (; t < tot; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % pitch);
        const int j = (int)((t / pitch) % (plane / pitch));
        if (i >= i0 && i < i1 && j >= j0 && j < j1) acc += np[mat[(long long)kl0 * plane + t]];
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// flags[kl] = 1 if plane kl of the material array holds an id with poles (ids >= first_disp)
__global__ void plane_pole_kernel(const uint8_t *mat, long long plane, int first_disp, int *flags) {
    const uint8_t *a = mat + (long long)blockIdx.y * plane;
    int any = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane && !any; t += (long long)gridDim.x * blockDim.x) any |= (a[t] >= first_disp);
    if (__syncthreads_or(any) && threadIdx.x == 0) flags[blockIdx.y] = 1;
}

__global__ void present_kernel(const uint8_t *a, long long n, unsigned *hist) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) sh[a[t]] = 1;
    __syncthreads();
    if (sh[threadIdx.x]) hist[threadIdx.x] = 1;
}

static int count_box(sj_sim *s, int i0, int i1, int j0, int j1, int k0, int k1, double *res) {
    *res = 0;
    k0 = std::max(k0, s->kz0); k1 = std::min(k1, s->kz1);
    if (k0 >= k1) return 0;
    unsigned long long *cnt; int rc = alloc_zero(s, (void **)&cnt, 8); if (rc) return rc;
    for (int c = 0; c < 3; ++c)
        count_pole_points<<<296, 256, 0, s->stream>>>(s->mat[c], s->mt_np, s->plane, s->pitch, i0, i1, j0, j1,
                                                       k0 - s->kz0 + 1, k1 - s->kz0 + 1, cnt);
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, cnt, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(cnt);
    *res = (double)h;
    return 0;
}

void sj_interior_geom(const sj_sim *s, int k_begin, int k_end, IntGeom &g, dim3 &grd) {
    const int V = s->prec == SJ_F64 ? 2 : 4;
    g.i_lo = s->lo[0]; g.i_hi = s->hi[0]; g.j_lo = s->lo[1]; g.j_hi = s->hi[1];
    g.k_lo = std::max(s->lo[2], s->kz0); g.k_hi = std::min(s->hi[2], s->kz1);
    g.zchunk = s->int_zchunk; g.flags = NULL;
    const int ni = std::max(g.i_hi - g.i_lo, 0), nj = std::max(g.j_hi - g.j_lo, 0);
    const int tw = s->int_lx * V, th = (32 / s->int_lx) * 8;
    const int kb = std::max(k_begin, g.k_lo), ke = std::min(k_end, g.k_hi);
    g.c0 = 0; g.nzc = 0;
    if (kb < ke) { g.c0 = (kb - g.k_lo) / g.zchunk; g.nzc = (ke - g.k_lo + g.zchunk - 1) / g.zchunk - g.c0; }
    grd = dim3((ni + tw - 1) / tw, (nj + th - 1) / th, std::max(g.nzc, 0) * s->g.n_sets);
}

// Material class of every work item, split along z (used by the register kernels' lists and by the TMA schedules).
// class of a tile plane from its flag word (plane_flags_kernel): mixed -> 1; one material without poles -> 0; one
// material with 1 or 2 poles -> 2 / 3 (SJ_NO_UNI=1 sends those through the general path too).
// Every (tile, z-chunk) item is cut along z into runs of planes that hold one material (fast paths: no material
// bytes, coefficients in registers) and runs that do not (general path).  Material interfaces in these scenes are
// mostly horizontal, so a chunk that straddles one is uniform above and below it.  Uniform runs shorter than
// `min_run` planes are not worth a block of their own and join the neighbouring general run.  The arithmetic of
// the paths is identical, so the split changes nothing but speed.  SJ_ZSPLIT=0: classify whole items.
int sj_classify_items(sj_sim *s, const std::vector<WorkItem> &items, int tile_w, int tile_h, std::vector<WorkItem> (&out)[4]) {
    static const bool uni_on = getenv("SJ_NO_UNI") == NULL;
    auto tile_class = [&](unsigned f) -> int {
        if (f & 1u) return 1;
        if (!(f & 2u)) return 0;
        const int np = s->mats_sorted[f >> 8].n_poles;
        return (uni_on && np >= 1 && np <= 2 && s->n_slots <= 2) ? 1 + np : 1;
    };
    static const int min_run = getenv("SJ_ZSPLIT") ? atoi(getenv("SJ_ZSPLIT")) : 3;
    const int ZMAX = 64;
    if (items.empty()) return 0;
    WorkItem *ditems; unsigned *df;
    CK(cudaMalloc((void **)&ditems, items.size() * sizeof(WorkItem)));
    CK(cudaMalloc((void **)&df, items.size() * ZMAX * sizeof(unsigned)));
    CK(cudaMemcpyAsync(ditems, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, s->stream));
    plane_flags_kernel<<<(unsigned)items.size(), 256, 0, s->stream>>>(s->mat[0], s->mat[1], s->mat[2], ditems, tile_w, tile_h,
        s->g.n[0], s->g.n[1], s->pitch, s->plane, s->kz0, s->first_disp, df, ZMAX);
    std::vector<unsigned> fl(items.size() * ZMAX);
    CK(cudaMemcpyAsync(fl.data(), df, fl.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaGetLastError());
    cudaFree(ditems); cudaFree(df);
    struct Seg { int kb, ke; unsigned key; };       // key 1 = general, else the flag word of a uniform run
    for (size_t n = 0; n < items.size(); ++n) {
        const WorkItem &it = items[n];
        const int nz = it.ke - it.kb;
        std::vector<Seg> seg;
        if (nz > ZMAX || min_run <= 0) {            // whole item: uniform only if every plane agrees
            unsigned key = nz > 0 && nz <= ZMAX ? fl[n * ZMAX] : 1u;
            if (tile_class(key) == 1) key = 1u;
            for (int kk = 1; kk < nz && nz <= ZMAX && key != 1u; ++kk) if (fl[n * ZMAX + kk] != key) key = 1u;
            seg.push_back(Seg{it.kb, it.ke, key});
        } else {
            for (int kk = 0; kk < nz; ++kk) {
                unsigned key = fl[n * ZMAX + kk];
                if (tile_class(key) == 1) key = 1u;
                if (!seg.empty() && seg.back().key == key) seg.back().ke = it.kb + kk + 1;
                else seg.push_back(Seg{it.kb + kk, it.kb + kk + 1, key});
            }
            if (seg.size() > 1) {
                for (size_t q = 0; q < seg.size(); ++q) if (seg[q].key != 1u && seg[q].ke - seg[q].kb < min_run) seg[q].key = 1u;
                std::vector<Seg> merged;
                for (size_t q = 0; q < seg.size(); ++q) {
                    if (!merged.empty() && merged.back().key == seg[q].key) merged.back().ke = seg[q].ke;
                    else merged.push_back(seg[q]);
                }
                seg.swap(merged);
            }
        }
        for (size_t q = 0; q < seg.size(); ++q) {
            WorkItem w = it;
            w.kb = seg[q].kb; w.ke = seg[q].ke;
            w.mat = seg[q].key == 1u ? 0 : (int)(seg[q].key >> 8);
            out[seg[q].key == 1u ? 1 : tile_class(seg[q].key)].push_back(w);
        }
    }
    return 0;
}

int sj_finish_materials(sj_sim *s) {
    graph_drop(s);                      // work lists and material tables are rebuilt below
    // sort the table: non-dispersive materials first, so "has poles" is a compare, not a load
    const int nm = (int)s->mats.size();
    std::vector<int> order;
    for (int np = 0; np <= SJ_MAX_POLES; ++np) {
        if (np == 1) s->first_disp = (int)order.size();
        for (int m = 0; m < nm; ++m) if (s->mats[m].n_poles == np) order.push_back(m);
    }
    uint8_t lut[256];
    for (int i = 0; i < 256; ++i) { lut[i] = 0; s->lut_inv[i] = (uint8_t)i; }
    bool ident = true;
    std::vector<sj_material> sorted(nm);
    for (int nw = 0; nw < nm; ++nw) { sorted[nw] = s->mats[order[nw]]; lut[order[nw]] = (uint8_t)nw; s->lut_inv[nw] = (uint8_t)order[nw]; ident &= (order[nw] == nw); }
    s->mats_sorted = sorted;
    {   // every index byte must name a table entry (an unknown id would be remapped to material 0 silently)
        unsigned *dh; CK(cudaMalloc((void **)&dh, 256 * sizeof(unsigned))); CK(cudaMemset(dh, 0, 256 * sizeof(unsigned)));
        CK(cudaDeviceSynchronize());
        for (int c = 0; c < 3; ++c)
            present_kernel<<<296, 256, 0, s->stream>>>(s->mat[c] + s->plane, s->plane * (s->nzl - 2), dh);
        unsigned hh[256];
        CK(cudaMemcpyAsync(hh, dh, sizeof hh, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        cudaFree(dh);
        for (int i = nm; i < 256; ++i) if (hh[i]) return fail(s, SJ_ERR_ARG, "material index outside the material table");
    }
    if (!ident) {
        uint8_t *dl; CK(cudaMalloc((void **)&dl, 256)); CK(cudaMemcpy(dl, lut, 256, cudaMemcpyHostToDevice));
        for (int c = 0; c < 3; ++c)
            remap_bytes<<<(unsigned)((s->set_stride + 255) / 256), 256, 0, s->stream>>>(s->mat[c], s->set_stride, dl);
        CK(cudaStreamSynchronize(s->stream));
        cudaFree(dl);
    }
    // which material ids occur at all (the rasterizer's table has 2^regions entries, most unused)
    {
        unsigned *dh; CK(cudaMalloc((void **)&dh, 256 * sizeof(unsigned))); CK(cudaMemset(dh, 0, 256 * sizeof(unsigned)));
        for (int c = 0; c < 3; ++c)
            present_kernel<<<296, 256, 0, s->stream>>>(s->mat[c] + s->plane, s->plane * (s->nzl - 2), dh);
        unsigned hh[256];
        CK(cudaMemcpyAsync(hh, dh, sizeof hh, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        cudaFree(dh);
        for (int i = 0; i < 256; ++i) s->present[i] = hh[i] != 0;
    }
    {   // local planes that hold a pole material (any E component): only those get polarisation storage
        int *df; CK(cudaMalloc((void **)&df, s->nzl * sizeof(int))); CK(cudaMemset(df, 0, s->nzl * sizeof(int)));
        CK(cudaDeviceSynchronize());
        for (int c = 0; c < 3; ++c)
            plane_pole_kernel<<<dim3(8, s->nzl), 256, 0, s->stream>>>(s->mat[c], s->plane, s->first_disp, df);
        std::vector<int> hf(s->nzl);
        CK(cudaMemcpyAsync(hf.data(), df, hf.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        cudaFree(df);
        int a = -1, b = -1;
        for (int kl = 1; kl < s->nzl - 1; ++kl) if (hf[kl]) { if (a < 0) a = kl; b = kl; }
        static const bool dense = getenv("SJ_P_DENSE") != NULL;
        if (dense) { a = 0; b = s->nzl - 1; }
        s->p_k0 = a < 0 ? 1 : a; s->p_nzp = a < 0 ? 1 : b - a + 1;
    }
    int rc = upload_material_table(s); if (rc) return rc;
    rc = count_box(s, 0, s->g.n[0] + 1, 0, s->g.n[1] + 1, s->kz0, s->kz1, &s->pole_points); if (rc) return rc;
    rc = count_box(s, s->lo[0], s->hi[0], s->lo[1], s->hi[1], s->lo[2], s->hi[2], &s->pole_points_int); if (rc) return rc;
    // block-uniform material flags for the E kernels
    {
        const int V = s->prec == SJ_F64 ? 2 : 4;
        IntGeom g; dim3 grd; sj_interior_geom(s, s->kz0, s->kz1, g, grd);
        auto upload = [&](ItemList &L, const std::vector<WorkItem> &v) -> int {
            cudaFree(L.dev); L.dev = NULL; L.n = (int)v.size();
            CK(cudaMalloc((void **)&L.dev, std::max<size_t>(v.size(), 1) * sizeof(WorkItem)));
            if (!v.empty()) CK(cudaMemcpy(L.dev, v.data(), v.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
            return 0;
        };
        auto classify = [&](const std::vector<WorkItem> &items, int tile_w, int tile_h, std::vector<WorkItem> (&out)[4]) -> int {
            return sj_classify_items(s, items, tile_w, tile_h, out);
        };
        // interior: one item per (tile, chunk); the field sets share the materials, so classify once and replicate
        {
            const int tw = s->int_lx * V, th = (32 / s->int_lx) * 8;
            std::vector<WorkItem> items;
            for (int kc = 0; kc < g.nzc; ++kc)
                for (unsigned ty = 0; ty < grd.y; ++ty)
                    for (unsigned tx = 0; tx < grd.x; ++tx) {
                        const int i0 = g.i_lo + (int)tx * tw, j0 = g.j_lo + (int)ty * th;
                        WorkItem w = {-1, 0, i0, j0, g.k_lo + kc * g.zchunk, std::min(g.k_lo + (kc + 1) * g.zchunk, g.k_hi), 0, 0,
                                      std::min(i0 + tw, g.i_hi), std::min(j0 + th, g.j_hi)};
                        if (w.kb < w.ke) items.push_back(w);
                    }
            std::vector<WorkItem> one[4], lst[4];
            rc = classify(items, tw, th, one); if (rc) return rc;
            for (int a = 0; a < 4; ++a)
                for (int q = 0; q < s->g.n_sets; ++q)
                    for (size_t n = 0; n < one[a].size(); ++n) { WorkItem w = one[a][n]; w.set = q; lst[a].push_back(w); }
            for (int a = 0; a < 4; ++a) { rc = upload(s->il_int[a], lst[a]); if (rc) return rc; }
        }
        for (int f = 0; f < 2; ++f)
            for (int wn = 0; wn < 2; ++wn) {
                const int lxw = wn ? s->pml_lx_n : s->pml_lx;
                std::vector<WorkItem> l2[4];
                rc = classify(s->h_items[f][wn], lxw * s->pml_v, (32 / lxw) * 8, l2); if (rc) return rc;
                for (int a2 = 0; a2 < 4; ++a2) { rc = upload(s->il_pml[f][wn][a2], l2[a2]); if (rc) return rc; }
            }
    }
    rc = sj_tma_build_materials(s); if (rc) return rc;
    s->materials_set = true;
    return 0;
}

extern "C" int sj_set_materials(sj_sim *s, int32_t n_mat, const sj_material *mats, const uint8_t *ix, const uint8_t *iy,
                                const uint8_t *iz) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || n_mat < 1 || n_mat > SJ_MAX_MAT || !mats) return fail(s, SJ_ERR_ARG, "bad material table");
    for (int m = 0; m < n_mat; ++m)
        if (mats[m].n_poles < 0 || mats[m].n_poles > SJ_MAX_POLES || !(mats[m].eps_inf > 0))
            return fail(s, SJ_ERR_ARG, "material with bad eps_inf or pole count");
    s->mats.assign(mats, mats + n_mat);
    const uint8_t *src[3] = {ix, iy, iz};
    const int nx1 = s->g.n[0] + 1, ny1 = s->g.n[1] + 1;
    for (int c = 0; c < 3; ++c) {
        if (!src[c]) { CK(cudaMemset(s->mat[c], 0, (size_t)s->set_stride)); continue; }
        // owned planes plus halos (clipped to the global grid) from the dense global host array
        for (int kl = 0; kl < s->nzl; ++kl) {
            const int k = s->kz0 - 1 + kl;
            if (k < 0 || k > s->g.n[2]) continue;
            CK(cudaMemcpy2D(s->mat[c] + (size_t)kl * s->plane, s->pitch, src[c] + (size_t)k * nx1 * ny1, nx1, nx1, ny1,
                            cudaMemcpyHostToDevice));
        }
    }
    return sj_finish_materials(s);
}

extern "C" int sj_get_material_table(sj_sim *s, int32_t *n_mat, sj_material *out, int32_t cap) {
    if (!s || !n_mat) return SJ_ERR_ARG;
    *n_mat = (int32_t)s->mats.size();
    for (int i = 0; i < *n_mat && i < cap && out; ++i) out[i] = s->mats[i];
    return SJ_OK;
}

extern "C" int sj_get_region_masks(sj_sim *s, int comp, uint8_t *out) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || comp < 0 || comp > 2 || !out) return SJ_ERR_ARG;
    const uint8_t *src = s->masks[comp] ? s->masks[comp] : s->mat[comp];
    const int nx1 = s->g.n[0] + 1, ny1 = s->g.n[1] + 1;
    CK(cudaStreamSynchronize(s->stream));
    for (int k = s->kz0; k < s->kz1; ++k)
        CK(cudaMemcpy2D(out + (size_t)(k - s->kz0) * nx1 * ny1, nx1, src + (size_t)(k - s->kz0 + 1) * s->plane, s->pitch,
                        nx1, ny1, cudaMemcpyDeviceToHost));
    if (!s->masks[comp])
        for (size_t i = 0; i < (size_t)nx1 * ny1 * (s->kz1 - s->kz0); ++i) out[i] = s->lut_inv[out[i]];
    return SJ_OK;
}

// ---- sources -------------------------------------------------------------------------------
// gaussian_src_time_phase (reference src/disp.cpp:378-400) evaluated on the host in fp64
static std::complex<double> src_dipole(const HostSource &g, double time) {
    if (g.kind == 2) { double o[2] = {0, 0}; g.fn(g.fn_ctx, time, o); return std::complex<double>(o[0], o[1]); }
    if (g.kind == 1) {
        // meep::continuous_src_time::dipole (what the reference instantiates for CW_source, src/disp.cpp:618):
        // zero outside [start, end] (float compare), exp(-i w t)/(-i w), tanh turn-on/off when width != 0
        const float rtime = float(time);
        if (rtime < g.t_start || rtime > g.t_end) return 0.0;
        const std::complex<double> osc = std::polar(1.0, -g.omega * time) * std::complex<double>(g.amp_t_re, g.amp_t_im);
        if (g.width == 0.0) return osc;
        const double ts = (time - g.t_start) / g.width - g.slowness, te = (g.t_end - time) / g.width - g.slowness;
        return osc * (1.0 + tanh(ts)) * (1.0 + tanh(te)) * 0.25;
    }
    const double tt = time - g.peak;
    if (float(fabs(tt)) > g.cutoff) return 0.0;
    return exp(-tt * tt / (2 * g.width * g.width)) * std::polar(1.0, -g.omega * tt - g.phi) *
           std::complex<double>(g.amp_t_re, g.amp_t_im);
}

extern "C" double sj_last_source_time(const sj_sim *s) {
    double t = 0;
    for (const auto &g : s->srcs) t = std::max(t, g.kind == 2 ? g.last_time : g.kind == 1 ? g.t_end : (double)float(g.peak + g.cutoff));
    return t;
}

// meep fields::add_volume_source + loop_in_chunks end-point weights for one direction
static int source_axis(const sj_sim *s, int comp, int d, double lo, double hi, int &a0, int &a1, std::vector<double> &w,
                       bool &delta) {
    const double a = s->g.a, inva = s->inva;
    const int n = s->g.n[d];
    const int sh = (d == comp) ? 1 : 0, iyc = 1 - sh;
    const double yc = iyc * (0.5 * inva);
    const int is = 1 + 2 * (int)floor((lo + yc) * a - .5) - iyc;
    const int ie = 1 + 2 * (int)ceil((hi + yc) * a - .5) - iyc;
    const double w0 = 1. - lo * a + 0.5 * is, w1 = 1. + hi * a - 0.5 * ie;
    double s0, s1, e0, e1;
    delta = (lo == hi);
    if (ie >= is + 6) { s0 = w0 * w0 / 2; s1 = 1 - (1 - w0) * (1 - w0) / 2; e0 = w1 * w1 / 2; e1 = 1 - (1 - w1) * (1 - w1) / 2; }
    else if (ie == is + 4) { s0 = w0 * w0 / 2; s1 = 1 - (1 - w0) * (1 - w0) / 2 - (1 - w1) * (1 - w1) / 2; e0 = w1 * w1 / 2; e1 = s1; }
    else if (lo == hi) { s0 = w0; s1 = w1; e0 = w1; e1 = w0; }
    else if (ie == is + 2) { s0 = w0 * w0 / 2 - (1 - w1) * (1 - w1) / 2; e0 = w1 * w1 / 2 - (1 - w0) * (1 - w0) / 2; s1 = e0; e1 = s0; }
    else return -1;
    const int own_lo = sh ? 0 : 1, own_hi = n - 1;  // index n is zeroed by the metallic wall
    const int i0 = (is - sh) / 2, i1 = (ie - sh) / 2;
    a0 = std::max(i0, own_lo); a1 = std::min(i1, own_hi);
    w.clear();
    for (int i = a0; i <= a1; ++i) {
        const int h = 2 * i + sh;
        w.push_back(h == is ? s0 : h == is + 2 ? s1 : h == ie ? e0 : h == ie - 2 ? e1 : 1.0);
    }
    return 0;
}

// source descriptors live in device memory (the kernels read them on the source planes only)
template <typename T>
static int upload_sources(sj_sim *s) {
    SrcDev<T> h[SJ_MAX_SRC];
    memset(h, 0, sizeof h);
    for (size_t q = 0; q < s->srcs.size(); ++q) {
        const HostSource &g = s->srcs[q];
        h[q].comp = g.comp;
        for (int d = 0; d < 3; ++d) { h[q].lo[d] = g.lo[d]; h[q].hi[d] = g.hi[d]; h[q].w[d] = (const T *)s->srcw[q][d]; }
        h[q].amp_re = (T)g.amp_re; h[q].amp_im = (T)g.amp_im;
    }
    if (!s->src_dev) CK(cudaMalloc(&s->src_dev, sizeof(SrcDev<double>) * SJ_MAX_SRC));
    CK(cudaMemcpy(s->src_dev, h, sizeof h, cudaMemcpyHostToDevice));
    return SJ_OK;
}

// place a source volume on the grid (meep fields::add_volume_source weights) and register it
static int place_source(sj_sim *s, HostSource &g, const double lo[3], const double hi[3], double amp_re, double amp_im,
                        const double *set_phase) {
    std::complex<double> amp(amp_re, amp_im);
    for (int d = 0; d < 3; ++d) {
        bool delta;
        if (source_axis(s, g.comp, d, lo[d], hi[d], g.lo[d], g.hi[d], g.w[d], delta)) return fail(s, SJ_ERR_ARG, "bad source volume");
        if (delta) amp *= s->g.a;
    }
    g.amp_re = amp.real(); g.amp_im = amp.imag();
    if (set_phase) g.set_phase.assign(set_phase, set_phase + s->g.n_sets);
    const int q = (int)s->srcs.size();
    for (int d = 0; d < 3; ++d) {
        int rc = s->prec == SJ_F64 ? upload_vec<double>(s, g.w[d], &s->srcw[q][d]) : upload_vec<float>(s, g.w[d], &s->srcw[q][d]);
        if (rc) return rc;
    }
    s->srcs.push_back(g);
    s->drive_dirty = true;
    return s->prec == SJ_F64 ? upload_sources<double>(s) : upload_sources<float>(s);
}

extern "C" int sj_add_gaussian_source(sj_sim *s, int comp, const double lo[3], const double hi[3], double amp_re,
                                      double amp_im, double freq, double width, double phase, double t_start,
                                      double t_end, int integrated, const double *set_phase) {
    if (!s || comp < 0 || comp > 2 || !lo || !hi) return fail(s, SJ_ERR_ARG, "bad source arguments (E components only)");
    if ((int)s->srcs.size() >= SJ_MAX_SRC) return fail(s, SJ_ERR_ARG, "too many sources");
    HostSource g;
    g.comp = comp; g.integrated = integrated; g.kind = 0; g.t_start = g.t_end = g.slowness = 0;
    g.omega = 2 * M_PI * freq; g.width = width; g.phi = phase + M_PI;
    g.peak = 0.5 * (t_start + t_end); g.cutoff = (t_end - t_start) * 0.5;
    std::complex<double> at = 1.0 / std::complex<double>(0, -g.omega);
    g.amp_t_re = at.real(); g.amp_t_im = at.imag();
    while (exp(-g.cutoff * g.cutoff / (2 * g.width * g.width)) < 1e-100) g.cutoff *= 0.9;
    g.cutoff = float(g.cutoff);
    return place_source(s, g, lo, hi, amp_re, amp_im, set_phase);
}

extern "C" int sj_add_cw_source(sj_sim *s, int comp, const double lo[3], const double hi[3], double amp_re, double amp_im,
                                double freq, double width, double t_start, double t_end, double slowness, int integrated,
                                const double *set_phase) {
    if (!s || comp < 0 || comp > 2 || !lo || !hi) return fail(s, SJ_ERR_ARG, "bad source arguments (E components only)");
    if ((int)s->srcs.size() >= SJ_MAX_SRC) return fail(s, SJ_ERR_ARG, "too many sources");
    if (!(t_end >= t_start) || !(t_end < 1e300)) return fail(s, SJ_ERR_ARG, "a CW source needs a finite end time");
    HostSource g;
    g.comp = comp; g.integrated = integrated; g.kind = 1;
    g.omega = 2 * M_PI * freq; g.width = width; g.phi = 0; g.peak = 0; g.cutoff = 0;
    g.t_start = t_start; g.t_end = t_end; g.slowness = slowness;
    std::complex<double> at = 1.0 / std::complex<double>(0, -g.omega);
    g.amp_t_re = at.real(); g.amp_t_im = at.imag();
    return place_source(s, g, lo, hi, amp_re, amp_im, set_phase);
}

// fields.add_volume_source(c, src, volume, amp) for an arbitrary meep::src_time subclass: fn is its dipole(time)
extern "C" int sj_add_custom_source(sj_sim *s, int comp, const double lo[3], const double hi[3], double amp_re, double amp_im,
                                    sj_dipole_fn fn, void *ctx, double last_time, int integrated, const double *set_phase) {
    if (!s || comp < 0 || comp > 2 || !lo || !hi || !fn) return fail(s, SJ_ERR_ARG, "bad source arguments (E components only)");
    if ((int)s->srcs.size() >= SJ_MAX_SRC) return fail(s, SJ_ERR_ARG, "too many sources");
    HostSource g;
    g.comp = comp; g.integrated = integrated; g.kind = 2; g.fn = fn; g.fn_ctx = ctx; g.last_time = last_time;
    g.omega = g.width = g.phi = g.peak = g.cutoff = g.t_start = g.t_end = g.slowness = g.amp_t_re = g.amp_t_im = 0;
    return place_source(s, g, lo, hi, amp_re, amp_im, set_phase);
}

// drive table [step][src][set][2]: {S_n, dt*J_n}; S_n = Re/Im(amp * dipole(n dt)) for integrated
// sources (meep update_eh), dt*J_n = Re/Im(amp * dt * current(n dt + dt/2)) otherwise (step_source)
static int ensure_drive(sj_sim *s, long long upto) {
    if (!s->drive_dirty && upto + 2 <= s->drive_steps) return 0;
    const long long steps = std::max<long long>(upto + 2, 2 * s->drive_steps);
    const int ns = (int)s->srcs.size(), nq = s->g.n_sets;
    std::vector<double> tab((size_t)steps * std::max(ns, 1) * nq * 2, 0.0);
    for (long long n = 0; n < steps; ++n)
        for (int q = 0; q < ns; ++q) {
            const HostSource &g = s->srcs[q];
            const std::complex<double> amp(g.amp_re, g.amp_im);
            std::complex<double> S = 0.0, J = 0.0;
            // The E kernels advance E by chi (dD - dP - (S_{n+1} - S_n)); with E = 0 at step 0 that equals meep's
            // E = chi (D - P - S) only if S_0 = 0, so a source that is already on at t = 0 enters with S_0 = 0.
            if (g.integrated) { if (n > 0) S = src_dipole(g, n * s->dt); }
            else {
                const double tm = n * s->dt + 0.5 * s->dt;
                J = ((src_dipole(g, tm + s->dt) - src_dipole(g, tm)) / s->dt) * s->dt;
            }
            for (int e = 0; e < nq; ++e) {
                double *o = &tab[(((size_t)n * ns + q) * nq + e) * 2];
                if (!g.set_phase.empty()) {
                    const std::complex<double> ph = std::polar(1.0, -g.set_phase[e]);
                    o[0] = (amp * S * ph).real(); o[1] = (amp * J * ph).real();
                } else if (e == 0) { o[0] = (amp * S).real(); o[1] = (amp * J).real(); }
                else if (e == 1) { o[0] = (amp * S).imag(); o[1] = (amp * J).imag(); }
            }
        }
    CK(cudaStreamSynchronize(s->stream));
    graph_drop(s);                      // the captured kernels hold the old table pointer
    cudaFree(s->drive); s->drive = NULL;
    int rc = s->prec == SJ_F64 ? upload_vec<double>(s, tab, &s->drive) : upload_vec<float>(s, tab, &s->drive);
    if (rc) return rc;
    s->h2d_bytes += (double)tab.size() * s->esz;
    s->drive_steps = steps; s->drive_dirty = false;
    return 0;
}

// ---- monitors --------------------------------------------------------------------------------
// meep grid_volume::interpolate: linear weights between the two bracketing Yee points per direction.
// Returns whether this slab owns the point (idx8 stays -1 otherwise).
static bool interp_stencil(const sj_sim *s, int comp, const double *xyz, long long *idx8, double *w8) {
    const int cdir = comp % 3; const bool isH = comp >= 3;
    int mid[3], sh[3]; double dv[3];
    for (int d = 0; d < 3; ++d) {
        sh[d] = isH ? (d != cdir) : (d == cdir);
        const double pc = xyz[d];
        const double p = (pc - sh[d] * (0.5 * s->inva)) * s->g.a;
        mid[d] = ((int)floor(p)) * 2 + 1 + sh[d];
        const double midv = mid[d] * (0.5 * s->inva);
        dv[d] = (pc - midv) * (2 * s->g.a);
    }
    // the rank owning the lower bracketing z plane evaluates the whole stencil (upper plane may be its halo)
    const int klow = (mid[2] - 1 - sh[2]) / 2;
    const int kown = std::min(std::max(klow, 0), s->g.n[2]);
    const bool mine = (kown >= s->kz0 && kown < s->kz1);
    for (int q = 0; q < 8; ++q) { idx8[q] = -1; w8[q] = 0.0; }
    if (!mine) return false;
    for (int q = 0; q < 8; ++q) {
        double wt = 1.0; long long lin = 0; bool ok = true;
        for (int d = 0; d < 3; ++d) {
            const int up = (q >> d) & 1;
            const int h = mid[d] + (up ? 1 : -1);
            wt *= up ? 0.5 * (1.0 + dv[d]) : 0.5 * (1.0 - dv[d]);
            const int i = (h - sh[d]) / 2;
            if (h - sh[d] < 0 || i > s->g.n[d]) { ok = false; continue; }
            if (d == 0) lin += i; else if (d == 1) lin += (long long)i * s->pitch;
            else { const int kl = i - s->kz0 + 1; if (kl < 0 || kl >= s->nzl) ok = false; else lin += (long long)kl * s->plane; }
        }
        if (wt < 0.0) wt = 0.0;
        if (!ok) wt = 0.0;
        idx8[q] = ok ? lin : 0;   // >= 0 marks "owned"; weight 0 drops the point
        w8[q] = wt;
    }
    return true;
}

extern "C" int sj_add_monitors(sj_sim *s, int comp, int32_t n, const double *xyz) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || comp < 0 || comp > 5 || n < 0 || (n && !xyz)) return fail(s, SJ_ERR_ARG, "bad monitor arguments");
    if (s->n_mon) return fail(s, SJ_ERR_STATE, "monitors already added");
    if (comp >= 3 && s->kz1 < s->g.n[2] + 1)
        return fail(s, SJ_ERR_UNSUPPORTED, "H-component monitors in a slab with an upper neighbour: the upper H halo is never exchanged");
    s->n_mon = n; s->mon_comp = comp;
    s->mon_xyz.assign(xyz, xyz + 3 * (size_t)n);
    std::vector<long long> idx((size_t)n * 8, -1);
    std::vector<double> w((size_t)n * 8, 0.0);
    s->mon_owned.assign(n, 0);
    for (int m = 0; m < n; ++m) s->mon_owned[m] = interp_stencil(s, comp, xyz + 3 * m, &idx[(size_t)m * 8], &w[(size_t)m * 8]);
    CK(cudaMalloc((void **)&s->mon_idx, std::max<size_t>(idx.size(), 1) * sizeof(long long)));
    CK(cudaMalloc((void **)&s->mon_w, std::max<size_t>(w.size(), 1) * sizeof(double)));
    if (n) {
        CK(cudaMemcpy(s->mon_idx, idx.data(), idx.size() * sizeof(long long), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s->mon_w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return SJ_OK;
}

static int ensure_series(sj_sim *s, int need) {
    if (need <= s->series_cap || s->n_mon == 0) return 0;
    const int cap = std::max(need, 2 * s->series_cap);
    double *nw;
    const size_t row = (size_t)s->n_mon * s->g.n_sets;
    // sj_sample / sj_pass take caller streams that are not ordered with s->stream: everything queued anywhere on the device
    // must have written the old buffer before it is copied and freed
    CK(cudaDeviceSynchronize());
    CK(cudaMalloc((void **)&nw, cap * row * sizeof(double)));
    CK(cudaMemsetAsync(nw, 0, cap * row * sizeof(double), s->stream));
    if (s->series) CK(cudaMemcpyAsync(nw, s->series, (size_t)s->n_samples * row * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(s->series);
    s->series = nw; s->series_cap = cap;
    return 0;
}

// ---- kernel parameter assembly + launches ------------------------------------------------------
static bool tma_step(const sj_sim *s) { return s->tma.mode == 3 && s->n_slots <= 2 && !s->trace_reg; }

// tick: the E-pass kernel of a full step also advances the device step counter (TMA path only; see tma_step)
static int do_pass(sj_sim *s, int which, int k0, int k1, cudaStream_t st, bool tick = false) {
    // whole-slab passes go through the TMA-staged column kernels (sj_tma.cuh); plane-restricted ones and scenes with more
    // than two pole slots through the register kernels
    if (((s->tma.mode >> which) & 1) && k0 <= s->kz0 && k1 >= s->kz1 && (which == 0 || s->n_slots <= 2) && !s->trace_reg)
        { const bool tk = tick && tma_step(s); return s->prec == SJ_F64 ? sj_tma_pass_f64(s, which, st, tk) : sj_tma_pass_f32(s, which, st, tk); }
    return s->prec == SJ_F64 ? sj_launch_pass_f64(s, which, k0, k1, st) : sj_launch_pass_f32(s, which, k0, k1, st);
}

static int do_sample(sj_sim *s, cudaStream_t st, long long base_step, int base_cursor, int span) {
    if (s->n_mon == 0) return 0;
    MonDev m; m.n_mon = s->n_mon; m.comp = s->mon_comp; m.idx = s->mon_idx; m.w = s->mon_w; m.series = s->series; m.flags = s->flags;
    const int nt = s->n_mon * s->g.n_sets;
    // a stencil may reach into the upper halo: wait until the upper slab's bottom E plane of the last step has landed
    m.wait_flag = s->peer_up.F ? s->sync_dev : NULL;
    if (s->prec == SJ_F64) { KParams<double> p; fill_params(s, p); sample_monitors<double><<<(nt + 127) / 128, 128, 0, st>>>(p, m, base_step, base_cursor, span); }
    else { KParams<float> p; fill_params(s, p); sample_monitors<float><<<(nt + 127) / 128, 128, 0, st>>>(p, m, base_step, base_cursor, span); }
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

// fields.get_field(c, loc) at arbitrary points, now: the monitors' interpolation kernel on a temporary point list
extern "C" int sj_sample_at(sj_sim *s, int comp, int32_t n, const double *xyz, double *out) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || comp < 0 || comp > 5 || n < 0 || (n && (!xyz || !out))) return fail(s, SJ_ERR_ARG, "bad sample arguments");
    if (n == 0) return SJ_OK;
    std::vector<long long> idx((size_t)n * 8, -1);
    std::vector<double> w((size_t)n * 8, 0.0);
    for (int m = 0; m < n; ++m) interp_stencil(s, comp, xyz + 3 * m, &idx[(size_t)m * 8], &w[(size_t)m * 8]);
    const size_t nt = (size_t)n * s->g.n_sets;
    long long *didx = NULL; double *dw = NULL, *dout = NULL; int *dflag = NULL;
    CK(cudaMalloc((void **)&didx, idx.size() * sizeof(long long)));
    CK(cudaMalloc((void **)&dw, w.size() * sizeof(double)));
    CK(cudaMalloc((void **)&dout, nt * sizeof(double)));
    CK(cudaMalloc((void **)&dflag, sizeof(int)));
    CK(cudaMemcpyAsync(didx, idx.data(), idx.size() * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(dw, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemsetAsync(dflag, 0, sizeof(int), s->stream));
    MonDev m; m.n_mon = n; m.comp = comp; m.idx = didx; m.w = dw; m.series = dout; m.flags = dflag;
    m.wait_flag = s->peer_up.F ? s->sync_dev : NULL;
    // cursor = 0: base_step is the current step and the span is never reached
    if (s->prec == SJ_F64) { KParams<double> p; fill_params(s, p); sample_monitors<double><<<(unsigned)((nt + 127) / 128), 128, 0, s->stream>>>(p, m, s->steps_done, 0, 1 << 30); }
    else { KParams<float> p; fill_params(s, p); sample_monitors<float><<<(unsigned)((nt + 127) / 128), 128, 0, s->stream>>>(p, m, s->steps_done, 0, 1 << 30); }
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, nt * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(didx); cudaFree(dw); cudaFree(dout); cudaFree(dflag);
    return SJ_OK;
}

extern "C" int sj_pass(sj_sim *s, int which, int32_t k0, int32_t k1, void *stream) {
    if (s) cudaSetDevice(s->g.device);
    if (!s) return SJ_ERR_ARG;
    int rc = ensure_drive(s, s->steps_done + 1); if (rc) return rc;
    return do_pass(s, which, k0, k1, stream ? (cudaStream_t)stream : s->stream);
}
extern "C" int sj_tick(sj_sim *s, void *stream) {
    if (s) cudaSetDevice(s->g.device);
    if (!s) return SJ_ERR_ARG;
    tick_kernel<<<1, 1, 0, stream ? (cudaStream_t)stream : s->stream>>>(s->step_dev);
    s->steps_done++; s->launches++;
    CK(cudaGetLastError());
    return 0;
}
extern "C" int sj_sample(sj_sim *s, void *stream) {
    if (s) cudaSetDevice(s->g.device);
    if (!s) return SJ_ERR_ARG;
    int rc = ensure_series(s, s->n_samples + 1); if (rc) return rc;
    rc = do_sample(s, stream ? (cudaStream_t)stream : s->stream, s->steps_done, s->n_samples, 1); if (rc) return rc;
    s->n_samples++;
    return 0;
}

// One full time step on stream st.  A slab whose planes fit the L2 window takes one launch: the fused wavefront kernel
// (step_tma in csrc/sj_tma.cuh); otherwise the H-pass and the E-pass are a launch each.
static bool fused_step(const sj_sim *s) {
    static const bool on = !(getenv("SJ_TMA_FUSE") && atoi(getenv("SJ_TMA_FUSE")) == 0);
    return on && tma_step(s) && s->tma.wave > 0 && s->tma.f.n_items > 0;
}
static int do_step(sj_sim *s, cudaStream_t st) {
    if (fused_step(s)) return s->prec == SJ_F64 ? sj_tma_pass_f64(s, 2, st, true) : sj_tma_pass_f32(s, 2, st, true);
    int rc = do_pass(s, 0, s->kz0, s->kz1, st); if (rc) return rc;
    rc = do_pass(s, 1, s->kz0, s->kz1, st, true); if (rc) return rc;
    if (!tma_step(s)) { tick_kernel<<<1, 1, 0, st>>>(s->step_dev); s->launches++; }
    return 0;
}

// Capture one full step on the simulation's stream.  Every kernel argument is step-invariant (the drive table is
// indexed by the device-side step counter), so the same graph serves every step until a setter changes a pointer.
static int graph_build(sj_sim *s) {
    const long long l0 = s->launches;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    int rc = do_step(s, s->stream);
    cudaGraph_t g = NULL;
    const cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
    s->graph_launches = s->launches - l0; s->launches = l0;
    // a failed capture is not an error of the run: fall back to direct launches (graph_launches < 0 = disabled)
    if (rc || ce != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); cudaGetLastError(); s->graph_launches = -1; return 0; }
    const cudaError_t ci = cudaGraphInstantiate(&s->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (ci != cudaSuccess) { cudaGetLastError(); s->step_graph = NULL; s->graph_launches = -1; }
    return 0;
}

// steps [i0, i1) of a run whose sampling cadence counts from step 0 of the run (reference src/disp.cpp:719-741):
// sample when i % save_span == 0, then step.  st: the stream the work is enqueued on.
static int run_range(sj_sim *s, int64_t i0, int64_t i1, int32_t save_span, long long base_step, int base_cursor, cudaStream_t st) {
    static const bool graph_env = getenv("SJ_NO_GRAPH") == NULL;
    const bool use_graph = graph_env && !s->trace_on;
    int rc;
    for (int64_t i = i0; i < i1; ++i) {
        if (i % save_span == 0) { rc = do_sample(s, st, base_step, base_cursor, save_span); if (rc) return rc; s->n_samples++; }
        if (use_graph && s->graph_launches >= 0) {
            if (!s->step_graph) { rc = graph_build(s); if (rc) return rc; }
            if (s->step_graph) {
                CK(cudaGraphLaunch(s->step_graph, st));
                s->steps_done++; s->launches += s->graph_launches;
                continue;
            }
        }
        rc = do_step(s, st); if (rc) return rc;
        s->steps_done++;
    }
    CK(cudaGetLastError());
    return SJ_OK;
}

// bound_geom::run loop body (reference src/disp.cpp:719-741)
extern "C" int sj_run(sj_sim *s, int64_t n_steps, int32_t save_span) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || n_steps < 0) return fail(s, SJ_ERR_ARG, "bad step count");
    if (save_span <= 0) save_span = 1;
    int rc = ensure_drive(s, s->steps_done + n_steps); if (rc) return rc;
    const int n_new = (int)((n_steps + save_span - 1) / save_span);
    rc = ensure_series(s, s->n_samples + n_new); if (rc) return rc;
    return run_range(s, 0, n_steps, save_span, s->steps_done, s->n_samples, s->stream);
}

// The same loop for the z-slabs of one simulation held by one process (one sj_sim per slab, stacked bottom to top and
// connected with sj_connect_local): the slabs order themselves through their mailbox flags on the device, so the host
// only has to keep every slab's queue fed -- the steps are enqueued round-robin in short chunks (a slab's kernels wait
// for its neighbours; a host thread blocked on one slab's full launch queue would starve the others).  Slabs that share
// a device are enqueued step by step on one stream, bottom slab first, which is the order their flags need.
extern "C" int sj_run_group(sj_sim **sims, int32_t n, int64_t n_steps, int32_t save_span) {
    if (!sims || n < 1 || n_steps < 0) return SJ_ERR_ARG;
    if (save_span <= 0) save_span = 1;
    bool shared = false;
    for (int a = 0; a < n; ++a) for (int b = a + 1; b < n; ++b) shared |= (sims[a]->g.device == sims[b]->g.device);
    std::vector<long long> base_step(n); std::vector<int> base_cursor(n);
    for (int a = 0; a < n; ++a) {
        sj_sim *s = sims[a];
        cudaSetDevice(s->g.device);
        int rc = ensure_drive(s, s->steps_done + n_steps); if (rc) return rc;
        rc = ensure_series(s, s->n_samples + (int)((n_steps + save_span - 1) / save_span)); if (rc) return rc;
        base_step[a] = s->steps_done; base_cursor[a] = s->n_samples;
        if (shared && a > 0) {          // everything goes through the bottom slab's stream: order it after this slab's set-up
            cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            cudaEventRecord(e, s->stream); cudaStreamWaitEvent(sims[0]->stream, e, 0); cudaEventDestroy(e);
        }
    }
    const int64_t chunk = shared ? 1 : 16;
    for (int64_t i0 = 0; i0 < n_steps; i0 += chunk)
        for (int a = 0; a < n; ++a) {
            sj_sim *s = sims[a];
            cudaSetDevice(s->g.device);
            int rc = run_range(s, i0, std::min(i0 + chunk, n_steps), save_span, base_step[a], base_cursor[a], shared ? sims[0]->stream : s->stream);
            if (rc) return rc;
        }
    if (shared)                         // later calls on the other slabs' own streams see the finished run
        for (int a = 1; a < n; ++a) {
            cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            cudaEventRecord(e, sims[0]->stream); cudaStreamWaitEvent(sims[a]->stream, e, 0); cudaEventDestroy(e);
        }
    return SJ_OK;
}

// ---- z-slab neighbours ---------------------------------------------------------------------------------------
struct PeerWire {              // what sj_export_peer writes into the caller's 256-byte buffer
    cudaIpcMemHandle_t f, sync;
    long long fcs, set_stride, plane;
    int nzl, kz0, kz1, n_sets, esz, magic;
};
static_assert(sizeof(PeerWire) <= 256, "sj_peer_handle too small");

extern "C" int sj_export_peer(sj_sim *s, sj_peer_handle *out) {
    if (!s || !out) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    PeerWire w; memset(&w, 0, sizeof w);
    CK(cudaIpcGetMemHandle(&w.f, s->F));
    CK(cudaIpcGetMemHandle(&w.sync, s->sync_dev));
    w.fcs = s->set_stride * s->g.n_sets; w.set_stride = s->set_stride; w.plane = s->plane;
    w.nzl = s->nzl; w.kz0 = s->kz0; w.kz1 = s->kz1; w.n_sets = s->g.n_sets; w.esz = (int)s->esz; w.magic = 0x534a5031;
    memset(out, 0, sizeof *out);
    memcpy(out, &w, sizeof w);
    return SJ_OK;
}

static int open_peer(sj_sim *s, const sj_peer_handle *h, bool upper, sj_sim::Peer &p) {
    PeerWire w; memcpy(&w, h, sizeof w);
    if (w.magic != 0x534a5031 || w.plane != s->plane || w.n_sets != s->g.n_sets || w.esz != (int)s->esz ||
        (upper ? w.kz0 != s->kz1 : w.kz1 != s->kz0))
        return fail(s, SJ_ERR_ARG, "peer handle does not describe the adjacent slab of the same simulation");
    CK(cudaIpcOpenMemHandle(&p.F, w.f, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&p.sync, w.sync, cudaIpcMemLazyEnablePeerAccess));
    p.fcs = w.fcs; p.set_stride = w.set_stride; p.nzl = w.nzl; p.ipc = true;
    return SJ_OK;
}

extern "C" int sj_connect_peers(sj_sim *s, const sj_peer_handle *lower, const sj_peer_handle *upper) {
    if (!s) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    if (s->steps_done) return fail(s, SJ_ERR_STATE, "connect the slabs before the first step");
    graph_drop(s);
    if (lower) { int rc = open_peer(s, lower, false, s->peer_down); if (rc) return rc; }
    if (upper) { int rc = open_peer(s, upper, true, s->peer_up); if (rc) return rc; }
    return SJ_OK;
}

extern "C" int sj_connect_local(sj_sim *lo, sj_sim *up) {
    if (!lo || !up || lo->kz1 != up->kz0 || lo->plane != up->plane || lo->esz != up->esz || lo->g.n_sets != up->g.n_sets)
        return fail(lo ? lo : up, SJ_ERR_ARG, "slabs are not stacked / incompatible");
    sj_sim *s = lo;
    if (lo->steps_done || up->steps_done) return fail(lo, SJ_ERR_STATE, "connect the slabs before the first step");
    if (lo->g.device != up->g.device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, lo->g.device, up->g.device));
        if (!can) return fail(lo, SJ_ERR_UNSUPPORTED, "no peer access between the two devices");
        cudaSetDevice(lo->g.device); cudaError_t e = cudaDeviceEnablePeerAccess(up->g.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaSetDevice(up->g.device); e = cudaDeviceEnablePeerAccess(lo->g.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
    }
    graph_drop(lo); graph_drop(up);
    lo->peer_up.F = up->F; lo->peer_up.sync = up->sync_dev; lo->peer_up.fcs = up->set_stride * up->g.n_sets;
    lo->peer_up.set_stride = up->set_stride; lo->peer_up.nzl = up->nzl; lo->peer_up.ipc = false;
    up->peer_down.F = lo->F; up->peer_down.sync = lo->sync_dev; up->peer_down.fcs = lo->set_stride * lo->g.n_sets;
    up->peer_down.set_stride = lo->set_stride; up->peer_down.nzl = lo->nzl; up->peer_down.ipc = false;
    return SJ_OK;
}

extern "C" int sj_run_timed(sj_sim *s, int64_t n_steps, int32_t save_span, double *ms) {
    if (!s || !ms) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    int rc = ensure_drive(s, s->steps_done + n_steps); if (rc) return rc;
    rc = ensure_series(s, s->n_samples + (int)((n_steps + std::max(save_span, 1) - 1) / std::max(save_span, 1))); if (rc) return rc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaEventRecord(e0, s->stream));
    rc = sj_run(s, n_steps, save_span);
    if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
    CK(cudaEventRecord(e1, s->stream));
    CK(cudaEventSynchronize(e1));
    float f = 0; CK(cudaEventElapsedTime(&f, e0, e1));
    *ms = f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return SJ_OK;
}

extern "C" int sj_profile_kernels(sj_sim *s, int32_t reps, double out[4]) {
    if (!s || !out || reps < 1) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    int rc = ensure_drive(s, s->steps_done + 1); if (rc) return rc;
    if (s->tma.mode == 3 && s->n_slots <= 2) {
        // the step is two persistent kernels (csrc/sj_tma.cuh): out[0] = the H-pass kernel, out[1] = the E-pass kernel, each
        // timed alone with CUDA events over `reps` back-to-back launches; the PML cells are inside them (out[2] = out[3] = 0)
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int which = 0; which < 2; ++which) {
            for (int rep = -2; rep < reps; ++rep) {
                if (rep == 0) CK(cudaEventRecord(e0, s->stream));
                rc = do_pass(s, which, s->kz0, s->kz1, s->stream); if (rc) return rc;
            }
            CK(cudaEventRecord(e1, s->stream));
            CK(cudaEventSynchronize(e1));
            float f = 0; CK(cudaEventElapsedTime(&f, e0, e1));
            out[which] = f / reps; out[2 + which] = 0.0;
        }
        if (fused_step(s)) {     // out[2] = the fused step kernel (both passes in one launch); it advances the step counter
            for (int rep = -2; rep < reps; ++rep) {
                if (rep == 0) CK(cudaEventRecord(e0, s->stream));
                rc = ensure_drive(s, s->steps_done + 2); if (rc) return rc;
                rc = do_step(s, s->stream); if (rc) return rc;
                s->steps_done++;
            }
            CK(cudaEventRecord(e1, s->stream));
            CK(cudaEventSynchronize(e1));
            float f = 0; CK(cudaEventElapsedTime(&f, e0, e1));
            out[2] = f / reps;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return SJ_OK;
    }
    return s->prec == SJ_F64 ? sj_profile_f64(s, reps, out) : sj_profile_f32(s, reps, out);
}

extern "C" int sj_get_counts(sj_sim *s, double out[6]) {
    if (!s || !out) return SJ_ERR_ARG;
    const int nk = s->kz1 - s->kz0;
    out[0] = (double)(s->g.n[0] + 1) * (s->g.n[1] + 1) * nk;
    const int kl = std::max(s->kz0, s->lo[2]), kh = std::min(s->kz1, s->hi[2]);
    out[1] = (double)std::max(0, s->hi[0] - s->lo[0]) * std::max(0, s->hi[1] - s->lo[1]) * std::max(0, kh - kl);
    out[2] = s->pml_cells;
    out[3] = s->pole_points;
    out[4] = s->pole_points_int;
    // true PML cells: any of the two half-pixel samples of the index has sigma != 0 in some direction
    double inter[3];
    for (int d = 0; d < 3; ++d) {
        int cnt = 0;
        const int a0 = d == 2 ? s->kz0 : 0, a1 = d == 2 ? s->kz1 : s->g.n[d] + 1;
        for (int i = a0; i < a1; ++i) if (s->sig[d][2 * i] == 0.0 && s->sig[d][2 * i + 1] == 0.0) ++cnt;
        inter[d] = cnt;
    }
    out[5] = out[0] - inter[0] * inter[1] * inter[2];
    return SJ_OK;
}

// per-plane sums of n_poles over the three E-component material arrays
__global__ void plane_pole_points_kernel(const uint8_t *m0, const uint8_t *m1, const uint8_t *m2, const int *np, long long plane,
                                         int pitch, int nx1, unsigned long long *out) {
    const uint8_t *a0 = m0 + (long long)(blockIdx.y + 1) * plane, *a1 = m1 + (long long)(blockIdx.y + 1) * plane, *a2 = m2 + (long long)(blockIdx.y + 1) * plane;
    unsigned long long acc = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane; t += (long long)gridDim.x * blockDim.x)
        if ((int)(t % pitch) < nx1) acc += np[a0[t]] + np[a1[t]] + np[a2[t]];
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + blockIdx.y, acc);
}

// Algorithmic bytes per step of every owned plane (the formula of sj_bytes_per_step, plane by plane): what a z-slab
// decomposition has to balance -- planes inside the substrate carry polarisation traffic, planes inside the z-PML the
// auxiliaries of every cell.  out: [kz1 - kz0].
extern "C" int sj_plane_costs(sj_sim *s, double *out) {
    if (!s || !out) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    const int nk = s->kz1 - s->kz0;
    unsigned long long *d; int rc = alloc_zero(s, (void **)&d, (size_t)nk * 8); if (rc) return rc;
    plane_pole_points_kernel<<<dim3(16, nk), 256, 0, s->stream>>>(s->mat[0], s->mat[1], s->mat[2], s->mt_np, s->plane, s->pitch,
                                                                   s->g.n[0] + 1, d);
    std::vector<unsigned long long> h(nk);
    CK(cudaMemcpyAsync(h.data(), d, (size_t)nk * 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(d);
    int in_xy[2];
    for (int dd = 0; dd < 2; ++dd) {
        int cnt = 0;
        for (int i = 0; i <= s->g.n[dd]; ++i) if (s->sig[dd][2 * i] == 0.0 && s->sig[dd][2 * i + 1] == 0.0) ++cnt;
        in_xy[dd] = cnt;
    }
    const double sz = (double)s->esz, per_plane = (double)s->g.n[0] * s->g.n[1], all_xy = (double)(s->g.n[0] + 1) * (s->g.n[1] + 1);
    for (int k = 0; k < nk; ++k) {
        const int kg = s->kz0 + k;
        const bool zin = s->sig[2][2 * kg] == 0.0 && s->sig[2][2 * kg + 1] == 0.0;
        const double pml = zin ? all_xy - (double)in_xy[0] * in_xy[1] : all_xy;
        out[k] = s->g.n_sets * (per_plane * (18.0 * sz + 3.0) + 3.0 * sz * (double)h[k] + 8.0 * sz * pml);
    }
    return SJ_OK;
}

extern "C" int sj_memory(const sj_sim *s, double out[6]) {
    if (!s || !out) return SJ_ERR_ARG;
    out[0] = 6.0 * (double)s->set_stride * s->g.n_sets * s->esz;                       // E, H
    out[1] = s->p_bytes;                                                               // polarisation (both parities)
    out[2] = s->pml_bytes;                                                             // UPML auxiliaries
    out[3] = 3.0 * (double)s->set_stride * (1.0 + (s->masks[0] ? 1.0 : 0.0));          // material index bytes (+ region masks)
    out[4] = out[0] + out[1] + out[2] + out[3];
    out[5] = (double)s->p_nzp;                                                         // planes with polarisation storage
    return SJ_OK;
}

extern "C" int sj_trace_step(sj_sim *s) {
    if (!s) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    int rc = ensure_drive(s, s->steps_done + 2); if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    s->trace_on = true; s->tr_ev.clear(); s->tr_name.clear();
    CK(cudaEventCreate(&s->tr_origin)); CK(cudaEventRecord(s->tr_origin, s->stream));
    rc = do_pass(s, 0, s->kz0, s->kz1, s->stream); if (rc) return rc;
    rc = do_pass(s, 1, s->kz0, s->kz1, s->stream, true); if (rc) return rc;
    if (!tma_step(s)) tick_kernel<<<1, 1, 0, s->stream>>>(s->step_dev);
    s->steps_done++;
    s->trace_on = false;
    CK(cudaStreamSynchronize(s->stream));
    for (int a = 0; a < s->n_aux; ++a) CK(cudaStreamSynchronize(s->aux[a]));
    for (size_t i = 0; i < s->tr_name.size(); ++i) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, s->tr_origin, s->tr_ev[2 * i]); cudaEventElapsedTime(&t1, s->tr_origin, s->tr_ev[2 * i + 1]);
        printf("trace %-28s start %8.1f us  end %8.1f us  dur %7.1f us\n", s->tr_name[i].c_str(), t0 * 1e3, t1 * 1e3, (t1 - t0) * 1e3);
    }
    for (auto e : s->tr_ev) cudaEventDestroy(e);
    s->tr_ev.clear(); s->tr_name.clear();
    return SJ_OK;
}

extern "C" int sj_sync(sj_sim *s) {
    if (s) cudaSetDevice(s->g.device);
    if (!s) return SJ_ERR_ARG;
    CK(cudaStreamSynchronize(s->stream));
    int fl[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(fl, s->flags, sizeof fl, cudaMemcpyDeviceToHost));
    if (fl[1]) return fail(s, SJ_ERR_STATE, "a z-slab neighbour did not deliver its boundary plane within 10 s");
    if (fl[0]) return fail(s, SJ_ERR_DIVERGED, "monitor sample exceeded 1000 or is not finite (divergence)");
    return SJ_OK;
}
extern "C" int64_t sj_steps_done(const sj_sim *s) { return s ? s->steps_done : -1; }
extern "C" int32_t sj_n_samples(const sj_sim *s) { return s ? s->n_samples : -1; }
extern "C" double sj_dt(const sj_sim *s) { return s ? s->dt : 0.0; }

extern "C" int sj_read_monitors(sj_sim *s, double *out) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || !out) return SJ_ERR_ARG;
    CK(cudaStreamSynchronize(s->stream));
    const size_t cnt = (size_t)s->n_samples * s->n_mon * s->g.n_sets;
    if (cnt) CK(cudaMemcpy(out, s->series, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    return SJ_OK;
}

// ---- spectra of the monitor series on the device (reference: fft() of src/data_utils.cpp:370-427 as called from
// save_field_times, disp.cpp:806).  X[m][q] = sum_{n < T} x_m[n] exp(-i (2 pi / N) n k(q)), T = 2^floor(log2 N) samples of
// the N taken, k(q) = q for q < T/2 and q - T above (FFT order).  One warp per (monitor, q): the lanes stride over n and
// the partial sums are combined with shuffles; the phase is reduced exactly in integers (n k mod N) before sincospi.
static __global__ void monitor_dft_kernel(const double *__restrict__ series, int n_mon, int n_sets, int set_re, int set_im, int T,
                                          int N, double *__restrict__ out) {
    const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_mon * T) return;
    const int mon = w / T, q = w % T;
    const long long k = (T > 1 && q >= T / 2) ? (long long)q - T : q;
    double ar = 0.0, ai = 0.0;
    for (int n = lane; n < T; n += 32) {
        long long r = (n * k) % N;
        if (r < 0) r += N;
        double sn, cs;
        sincospi(-2.0 * (double)r / (double)N, &sn, &cs);
        const double *x = series + ((long long)n * n_mon + mon) * n_sets;
        const double xr = x[set_re], xi = set_im >= 0 ? x[set_im] : 0.0;
        ar += xr * cs - xi * sn;
        ai += xr * sn + xi * cs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ar += __shfl_down_sync(0xffffffffu, ar, o); ai += __shfl_down_sync(0xffffffffu, ai, o); }
    if (lane == 0) { out[2 * (long long)w] = ar; out[2 * (long long)w + 1] = ai; }
}

extern "C" int sj_read_spectra(sj_sim *s, int32_t set_re, int32_t set_im, int32_t *n_freq, double *out) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || !n_freq) return SJ_ERR_ARG;
    if (set_re < 0 || set_re >= s->g.n_sets || set_im >= s->g.n_sets) return fail(s, SJ_ERR_ARG, "bad field set");
    const int N = s->n_samples;
    int T = 0;
    if (N > 0) T = 1 << (int)(log((double)N) / log(2.0));       // the reference's own expression (data_utils.cpp:417)
    *n_freq = T;
    if (!out || !T || !s->n_mon) return SJ_OK;
    double *dev = NULL;
    const size_t cnt = (size_t)s->n_mon * T * 2;
    CK(cudaMalloc((void **)&dev, cnt * sizeof(double)));
    const long long threads = (long long)s->n_mon * T * 32;
    monitor_dft_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s->stream>>>(s->series, s->n_mon, s->g.n_sets, set_re, set_im, T, N, dev);
    s->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, dev, cnt * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(s, SJ_ERR_CUDA, cudaGetErrorString(e));
    return SJ_OK;
}

extern "C" int sj_plane_ptr(sj_sim *s, int comp, int set, int32_t k, void **ptr, size_t *bytes) {
    if (!s || comp < 0 || comp > 5 || set < 0 || set >= s->g.n_sets || !ptr) return SJ_ERR_ARG;
    const int kl = k - s->kz0 + 1;
    if (kl < 0 || kl >= s->nzl) return fail(s, SJ_ERR_ARG, "plane outside local slab");
    char *base = (char *)(comp < 3 ? s->E[comp] : s->H[comp - 3]);
    *ptr = base + ((size_t)set * s->set_stride + (size_t)kl * s->plane) * s->esz;
    if (bytes) *bytes = (size_t)s->plane * s->esz;
    return SJ_OK;
}

extern "C" int sj_halo_exchange(sj_sim *lo, sj_sim *up, int which) {
    if (!lo || !up || lo->kz1 != up->kz0 || lo->plane != up->plane || lo->esz != up->esz || lo->g.n_sets != up->g.n_sets)
        return fail(lo ? lo : up, SJ_ERR_ARG, "slabs are not stacked / incompatible");
    sj_sim *s = lo;
    cudaSetDevice(which == 0 ? up->g.device : lo->g.device);
    sj_sim *src = which == 0 ? lo : up, *dst = which == 0 ? up : lo;
    const int k = which == 0 ? lo->kz1 - 1 : up->kz0;     // global plane that travels
    const size_t bytes = (size_t)lo->plane * lo->esz;
    CK(cudaEventRecord(src->ev_a, src->stream));
    CK(cudaStreamWaitEvent(dst->stream, src->ev_a, 0));
    for (int set = 0; set < lo->g.n_sets; ++set)
        for (int c = 0; c < 2; ++c) {
            void *ps, *pd;
            int rc = sj_plane_ptr(src, which == 0 ? 3 + c : c, set, k, &ps, NULL); if (rc) return rc;
            rc = sj_plane_ptr(dst, which == 0 ? 3 + c : c, set, k, &pd, NULL); if (rc) return rc;
            CK(cudaMemcpyPeerAsync(pd, dst->g.device, ps, src->g.device, bytes, dst->stream));
        }
    CK(cudaEventRecord(dst->ev_b, dst->stream));
    CK(cudaStreamWaitEvent(src->stream, dst->ev_b, 0));
    return SJ_OK;
}

extern "C" int sj_get_field(sj_sim *s, int comp, int set, double *out) {
    if (s) cudaSetDevice(s->g.device);
    if (!s || comp < 0 || comp > 5 || set < 0 || set >= s->g.n_sets || !out) return SJ_ERR_ARG;
    CK(cudaStreamSynchronize(s->stream));
    const int nx1 = s->g.n[0] + 1, ny1 = s->g.n[1] + 1, nk = s->kz1 - s->kz0;
    const char *base = (const char *)(comp < 3 ? s->E[comp] : s->H[comp - 3]) + ((size_t)set * s->set_stride + s->plane) * s->esz;
    if (s->prec == SJ_F64) {
        for (int k = 0; k < nk; ++k)
            CK(cudaMemcpy2D(out + (size_t)k * nx1 * ny1, nx1 * 8, base + (size_t)k * s->plane * 8, s->pitch * 8, nx1 * 8, ny1, cudaMemcpyDeviceToHost));
    } else {
        std::vector<float> tmp((size_t)nx1 * ny1);
        for (int k = 0; k < nk; ++k) {
            CK(cudaMemcpy2D(tmp.data(), nx1 * 4, base + (size_t)k * s->plane * 4, s->pitch * 4, nx1 * 4, ny1, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < tmp.size(); ++i) out[(size_t)k * nx1 * ny1 + i] = tmp[i];
        }
    }
    return SJ_OK;
}

extern "C" int sj_get_stats(const sj_sim *s, int64_t *launches, double *reserved) {
    if (!s) return SJ_ERR_ARG;
    if (launches) *launches = s->launches;
    if (reserved) *reserved = s->h2d_bytes;     // bytes of source drive table uploaded so far (host -> device)
    return SJ_OK;
}

// Algorithmic bytes per time step: every field array once per pass
// (18 s per cell), one material byte per E component, 3 s per pole per E-component point,
// and ~8 s per PML cell for the auxiliaries; all per field set.
extern "C" double sj_bytes_per_step(const sj_sim *s) {
    if (!s) return 0.0;
    const double cells = (double)s->g.n[0] * s->g.n[1] * (double)std::min(s->kz1 - s->kz0, s->g.n[2]);
    const double sz = (double)s->esz;
    double cnt[6]; sj_get_counts((sj_sim *)s, cnt);
    return s->g.n_sets * (cells * (18.0 * sz + 3.0) + 3.0 * sz * s->pole_points + 8.0 * sz * cnt[5]);
}
