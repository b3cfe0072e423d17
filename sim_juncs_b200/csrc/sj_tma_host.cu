// sj_tma_host.cu -- host side of the TMA-staged column kernels (sj_tma.cuh): tile shapes per region, CUtensorMap
// descriptors of the field / polarisation / PML-box allocations, work items and the static per-block schedules.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "sj_tma.cuh"

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = NULL;
    if (!fn) {
        void *p = NULL;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rank-3 map over an allocation stored [unit][row][x]: dim0 = pitch elements, dim1 = rows, dim2 = units (planes of all
// arrays and sets back to back); box = (bw, bh, 1).  Out-of-range elements are filled with zeros.
static int make_map(sj_sim *s, CUtensorMap *m, void *base, int pitch, int rows, long long units, long long unit_stride, int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { s->err = "cuTensorMapEncodeTiled not available from this driver"; return SJ_ERR_CUDA; }
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)std::max<long long>(units, 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * s->esz, (cuuint64_t)unit_stride * s->esz};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // L2 promotion off: tiles start on 32-byte sectors, not on 128- or 256-byte lines, and a promoted fetch would pull the
    // rest of every line it touches from DRAM (measured: +46 % read traffic with L2_256B)
    static const int promo = getenv("SJ_TMA_L2PROMO") ? atoi(getenv("SJ_TMA_L2PROMO")) : 0;
    const CUtensorMapL2promotion pr = promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    const CUresult r = fn(m, s->prec == SJ_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[256];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d): pitch %d rows %d units %lld box %d x %d", (int)r, pitch, rows, units, bw, bh);
        s->err = b;
        return SJ_ERR_CUDA;
    }
    return 0;
}

// tile shape for a w x h rectangle: nvx vectors per row, th rows; a thread owns one vector of up to max_sub rows (ths rows
// apart), so nvx * ceil(th / sub) <= NT threads.  Minimises the vectors a plane of the region makes the TMA engine read
// (three curl inputs with halo + three own arrays, tile overhang past the region included) per useful vector; ties go to
// the fuller thread block.  Tall tiles (sub = 2) halve the number of TMA operations per byte: the TMA unit of an SM takes
// ~170 cycles per box of up to ~18 rows whatever its size (scripts/microbench/tma_rate.cu), which at 4 KB per box is no
// more than the HBM rate.
static void pick_shape(int w, int h, int V, int NT, int max_sub, int &nvx_out, int &th_out, int &sub_out) {
    const int nv = (w + V - 1) / V;
    double best = 1e30;
    nvx_out = 1; th_out = 1; sub_out = 1;
    for (int sub = 1; sub <= max_sub; ++sub)
        for (int nvx = 1; nvx <= std::min(nv, 63); ++nvx) {
            const int tiles_x = (nv + nvx - 1) / nvx;
            if ((nv + tiles_x - 1) / tiles_x != nvx) continue;          // only balanced splits
            if (nvx < std::min(nv, 6)) continue;                         // rows of at least 96 bytes where the region allows
            if ((nvx & 1) && nvx != nv) continue;                        // tile rows are whole 32-byte sectors (DRAM granularity)
            for (int th = 1; th <= std::min(h, 255); ++th) {
                const int ths = (th + sub - 1) / sub;
                if (nvx * ths > NT) continue;
                if (sub > 1 && th < 2 * sub) continue;
                const int tiles_y = (h + th - 1) / th;
                if ((h + tiles_y - 1) / tiles_y != th) continue;
                const double read = (double)tiles_x * tiles_y * (3.0 * (nvx + 1) * (th + 1) + 3.0 * nvx * th) / ((double)nv * h);
                const double fill = (double)nvx * th / (NT * sub);
                const double score = read + 0.6 * (1.0 - fill);
                if (score < best) { best = score; nvx_out = nvx; th_out = th; sub_out = sub; }
            }
        }
}

static int shape_index(sj_sim *s, int nvx, int th, int sub, int V) {
    TmaState &t = s->tma;
    for (int i = 0; i < t.n_shapes; ++i) if (t.shapes[i].nvx == nvx && t.shapes[i].th == th && t.shapes[i].sub == sub) return i;
    if (t.n_shapes >= SJ_TMA_MAX_SHAPES) return -1;
    const int hs = ((nvx + 1) * (th + 1) * 16 + 127) / 128 * 128, os = (nvx * th * 16 + 127) / 128 * 128;
    t.shapes[t.n_shapes] = TShape{nvx, th, nvx * V, (nvx + 1) * V, sub, (th + sub - 1) / sub, hs, os};
    return t.n_shapes++;
}

static void free_list(TmaList &l) { cudaFree(l.items); cudaFree(l.first); l.items = NULL; l.first = NULL; l.n_items = 0; l.grid = 0; }

void sj_tma_free(sj_sim *s) {
    TmaState &t = s->tma;
    cudaFree(t.maps); t.maps = NULL;
    for (int g = 0; g < 2; ++g) { free_list(t.h[g]); for (int c = 0; c < 4; ++c) free_list(t.e[g][c]); }
    free_list(t.f); cudaFree(t.grp); t.grp = NULL; t.wave = 0;
}

// Static schedule: items sorted by cost, each given to the least-loaded block (LPT); a block then walks its items in
// that order.  cost = planes loaded (run + 1 partial plane).
// bytes one plane of the item stages in the ring (the cost unit of the schedule): 3 halo tiles of 20 B per thread and
// own tiles of 16 B per thread -- own fields, auxiliaries, 6 polarisation tiles per pole slot (E-pass)
static int item_weight(const sj_sim *s, const WorkItem &w, int which) {
    const bool general = w.box >= 0 && w.kind == 0;
    const int ns = which == 0 ? 0 : (w.pad == 0 ? 0 : w.pad == 1 ? std::max(s->n_slots, 1) : w.pad - 1);
    // edge / corner and mixed-material tiles are bound by instruction latency, not by their bytes: measured factors
    static const double wgen = getenv("SJ_TMA_WGEN") ? atof(getenv("SJ_TMA_WGEN")) : 1.0, wmix = getenv("SJ_TMA_WMIX") ? atof(getenv("SJ_TMA_WMIX")) : 1.0;
    const double f = (general ? wgen : 1.0) * ((which == 1 && w.pad == 1) ? wmix : 1.0);
    const TShape &sh = s->tma.shapes[w.shape & 0xff];
    return (int)(f * (3 * sh.hs + (3 + (general ? 6 : w.box >= 0 ? 1 : 0) + 6 * ns) * sh.os) / 64);
}

static int upload_schedule(sj_sim *s, std::vector<WorkItem> items, int which, int grid_cap, TmaList &out) {
    free_list(out);
    if (items.empty()) return 0;
    auto cost = [&](const WorkItem &w) { return (long long)(w.ke - w.kb + 1) * item_weight(s, w, which); };
    auto by_cost = [&](const WorkItem &a, const WorkItem &b) {
        const int ba = (a.shape & SJ_BND_FLAG) != 0, bb = (b.shape & SJ_BND_FLAG) != 0;
        return ba != bb ? ba > bb : cost(a) > cost(b);
    };
    std::stable_sort(items.begin(), items.end(), by_cost);
    {   // the last quarter of the work is cut into runs of at most `fine` planes: fine grains even out the end of the kernel
        static const int fine = getenv("SJ_TMA_FINE") ? atoi(getenv("SJ_TMA_FINE")) : 6;
        const size_t keep = items.size() - items.size() / 4;
        std::vector<WorkItem> cut(items.begin(), items.begin() + keep);
        for (size_t n = keep; n < items.size() && fine > 0; ++n) {
            const WorkItem &w = items[n];
            const int nz = w.ke - w.kb, parts = std::max(1, (nz + fine - 1) / fine), len = (nz + parts - 1) / parts;
            for (int kb = w.kb; kb < w.ke; kb += len) { WorkItem c = w; c.kb = kb; c.ke = std::min(kb + len, w.ke); cut.push_back(c); }
        }
        if (fine > 0) items.swap(cut);
        std::stable_sort(items.begin(), items.end(), by_cost);
    }
    // dynamic queue: the kernels' producers pull the next item with an atomic, heaviest first
    const int grid = std::max(1, std::min(grid_cap, (int)items.size()));
    const std::vector<WorkItem> &flat = items;
    const int zero[4] = {0, 0, 0, 0};
    CK(cudaMalloc((void **)&out.items, flat.size() * sizeof(WorkItem)));
    CK(cudaMalloc((void **)&out.first, sizeof zero));
    CK(cudaMemcpy(out.items, flat.data(), flat.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(out.first, zero, sizeof zero, cudaMemcpyHostToDevice));
    out.n_items = (int)flat.size(); out.grid = grid;
    return 0;
}

static std::vector<WorkItem> replicate_sets(const std::vector<WorkItem> &one, int n_sets) {
    std::vector<WorkItem> all;
    for (int q = 0; q < n_sets; ++q)
        for (WorkItem w : one) { w.set = q; all.push_back(w); }
    return all;
}

static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

int sj_tma_build_geometry(sj_sim *s) {
    TmaState &t = s->tma;
    t.n_shapes = 0; t.geo[0].clear(); t.geo[1].clear();
    if (!t.mode) return 0;
    int dev = 0, cc = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, dev);
    if (cc < 9 || !encode_fn()) { t.mode = 0; return 0; }          // no TMA on this device / driver: register kernels only
    const int V = s->prec == SJ_F64 ? 2 : 4;
    t.nt = SJ_TMA_NT;
    int n_sm_ = 148;
    cudaDeviceGetAttribute(&n_sm_, cudaDevAttrMultiProcessorCount, dev);
    // planes per item: about 16, fewer in a thin slab so that every SM still gets several items to pull (a slab of 23
    // planes cut into two runs per tile leaves 1.6 items per block and a third of the machine idle at the end)
    int zc_int = env_int("SJ_TMA_ZC", 0);
    if (zc_int <= 0) {
        const int V_ = s->prec == SJ_F64 ? 2 : 4;
        const long long tiles_xy = (long long)((s->g.n[0] + 1 + 26 * V_ - 1) / (26 * V_)) * ((s->g.n[1] + 1 + 15) / 16);
        const long long planes = (long long)(s->kz1 - s->kz0) * tiles_xy * s->g.n_sets;
        zc_int = (int)std::min<long long>(s->int_zchunk, std::max<long long>(4, planes / (5LL * n_sm_)));
    }
    const int sub_int = env_int("SJ_TMA_SUB", 2), sub_face = 1;     // (the kernels take tall tiles for interior items only)
    const int zc_gen = env_int("SJ_TMA_ZCG", 6);       // edge / corner tiles are bound by instruction latency: many short items
    struct Reg { int box, kind, i0, i1, j0, j1, k0, k1; };
    std::vector<Reg> regs;
    const int N1x = s->g.n[0] + 1, N1y = s->g.n[1] + 1;
    const int L[3] = {s->lo[0], s->lo[1], s->lo[2]}, Hh[3] = {s->hi[0], s->hi[1], s->hi[2]};
    regs.push_back(Reg{-1, 0, L[0], Hh[0], L[1], Hh[1], std::max(L[2], s->kz0), std::min(Hh[2], s->kz1)});
    for (size_t bi = 0; bi < s->boxes.size(); ++bi) {
        const sj_sim::Box &B = s->boxes[bi];
        regs.push_back(Reg{(int)bi, B.kind, B.lo[0], B.hi[0], B.lo[1], B.hi[1], B.lo[2], B.hi[2]});
    }
    std::vector<WorkItem> geo;
    for (const Reg &R : regs) {
        const int w = R.i1 - R.i0, h = R.j1 - R.j0, nz = R.k1 - R.k0;
        if (w <= 0 || h <= 0 || nz <= 0) continue;
        // tall tiles (two rows per thread) for the interior and, optionally, the single-sigma faces; the half-height shape
        // of a tall one is registered too: material classes whose plane load would not fit the ring twice take it
        const int max_sub = R.box < 0 ? sub_int : (R.kind != 0 ? sub_face : 1);
        int nvx, th, sub;
        pick_shape(w, h, V, t.nt, max_sub, nvx, th, sub);
        const int sh = shape_index(s, nvx, th, sub, V);
        if (sh < 0 || (sub > 1 && shape_index(s, nvx, (th + sub - 1) / sub, 1, V) < 0)) { t.mode = 0; return 0; }
        // planes per item: runs of about zc_int planes, evened out; thin regions (the z boxes) in one run
        const int zc_t = (R.box >= 0 && R.kind == 0) ? zc_gen : zc_int;
        const int nchunk = std::max(1, (nz + zc_t - 1) / zc_t), zc = (nz + nchunk - 1) / nchunk;
        const int tw = nvx * V;
        for (int kb = R.k0; kb < R.k1; kb += zc)
            for (int j0 = R.j0; j0 < R.j1; j0 += th)
                for (int i0 = R.i0; i0 < R.i1; i0 += tw) {
                    WorkItem wi = {R.box, 0, i0, j0, kb, std::min(kb + zc, R.k1), 0, R.kind, std::min(i0 + tw, R.i1), std::min(j0 + th, R.j1), sh, 0};
                    geo.push_back(wi);
                }
    }
    // A slab with a neighbour above (below) sends the top H plane (bottom E plane) of every step: that plane becomes
    // items of its own, flagged and first in the queue, so it is on its way while the rest of the slab is updated.
    t.n_bnd[0] = t.n_bnd[1] = 0;
    for (int which = 0; which < 2; ++which) {
        const bool has = which == 0 ? (s->kz1 < s->g.n[2] + 1) : (s->kz0 > 0);
        const int kb_ = which == 0 ? s->kz1 - 1 : s->kz0;
        for (const WorkItem &w : geo) {
            if (!has || !(w.kb <= kb_ && kb_ < w.ke)) { t.geo[which].push_back(w); continue; }
            WorkItem b = w; b.kb = kb_; b.ke = kb_ + 1; b.shape |= SJ_BND_FLAG;
            t.geo[which].push_back(b); t.n_bnd[which]++;
            if (w.kb < kb_) { WorkItem r = w; r.ke = kb_; t.geo[which].push_back(r); }
            if (kb_ + 1 < w.ke) { WorkItem r = w; r.kb = kb_ + 1; t.geo[which].push_back(r); }
        }
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    return upload_schedule(s, replicate_sets(t.geo[0], s->g.n_sets), 0, n_sm, t.h[0]);
}

// Fused step: both passes' items in one queue, cut along z into chunks of `wave` planes and ordered as a wavefront -- H(0), H(1),
// H(2), E(0), H(3), E(1), ... (heaviest first inside a group) -- so that an E-pass item finds what the H-pass items of its chunk
// wrote, and the E planes they read, in the L2.  A slab with z neighbours sweeps from its top chunk down instead: its top H
// plane (the upper slab's E-pass waits for it) is then the first item of the step and its bottom E plane (the lower slab's
// next H-pass waits for it) the last, which is the order the boundary-plane exchange needs.  H groups are labelled in queue
// order; an E-pass item carries the label of the last H group it depends on (bits 16.. of WorkItem::shape).  Only while a
// few chunks of both passes fit the L2 window; SJ_TMA_WAVE = planes per chunk (0: off).
static int build_fused(sj_sim *s, const std::vector<WorkItem> &e_items, int n_sm) {
    TmaState &t = s->tma;
    free_list(t.f); cudaFree(t.grp); t.grp = NULL; t.wave = 0; t.n_chunks = 0;
    const double plane_bytes = (double)s->plane * s->esz * s->g.n_sets * (18.0 + 6.0 * std::max(s->n_slots, 0));   // both passes, one plane
    // chunk length: measured on the bench workload (17 MB per plane of both passes): DRAM reads per step 2033 MB with two
    // launches, 1747 / 1465 / 1210 MB with 8 / 6 / 2-plane chunks -- but short chunks pay more per item (dependency waits,
    // partial planes, code switches) than the saved bytes are worth: 0.50 / 0.486 / 0.479 / 0.472 / 0.470 ms per step with
    // 3 / 4 / 5 / 8 / 10 planes.  Planes too large for the L2 to hold a chunk of both passes stay with two launches per step.
    int wave = env_int("SJ_TMA_WAVE", -1);
    if (wave < 0) { wave = std::min(8, (int)(140e6 / plane_bytes)); if (wave < 2) wave = 0; }
    if (wave <= 0) return 0;
    const bool down = t.n_bnd[0] || t.n_bnd[1];              // a slab with neighbours sweeps top-down
    const int nz = s->kz1 - s->kz0, nch = (nz + wave - 1) / wave;
    std::vector<std::vector<WorkItem>> grp[2];
    grp[0].resize(nch); grp[1].resize(nch);
    auto label = [&](int c) { return down ? nch - 1 - c : c; };      // position of chunk c in the sweep
    // the last two groups of the step are cut into runs of at most `fine` planes: fine grains even out the end of the launch
    const int fine = std::max(1, env_int("SJ_TMA_FINE", 3));
    auto cut = [&](const std::vector<WorkItem> &in, int pass) {
        for (const WorkItem &w : in)
            for (int kb = w.kb; kb < w.ke;) {
                const int c = (kb - s->kz0) / wave;
                int ke = std::min(w.ke, s->kz0 + (c + 1) * wave);
                if (pass == 1 && label(c) >= nch - 2) ke = std::min(ke, kb + fine);
                // an E-pass item of chunk c needs the H-pass items of the chunks c - 1 and c: the later of the two in the sweep
                const int lab = pass == 0 ? label(c) : std::max(label(c), c > 0 ? label(c - 1) : 0);
                for (int q = 0; q < s->g.n_sets; ++q) {
                    WorkItem x = w; x.kb = kb; x.ke = ke; x.set = q;
                    x.shape = (w.shape & (0xff | SJ_BND_FLAG)) | (pass ? SJ_EPASS_FLAG : 0) | (lab << 16);
                    grp[pass][label(c)].push_back(x);
                }
                kb = ke;
            }
    };
    cut(t.geo[0], 0); cut(e_items, 1);
    auto cost = [&](const WorkItem &w) { return (long long)(w.ke - w.kb + 1) * item_weight(s, w, (w.shape & SJ_EPASS_FLAG) ? 1 : 0); };
    std::vector<WorkItem> flat;
    std::vector<int> need(2 * nch + 1, 0);
    auto emit = [&](int pass, int g) {
        std::vector<WorkItem> &v = grp[pass][g];
        std::stable_sort(v.begin(), v.end(), [&](const WorkItem &a, const WorkItem &b) {
            const int ba = (a.shape & SJ_BND_FLAG) != 0, bb = (b.shape & SJ_BND_FLAG) != 0;
            return ba != bb ? (pass == 0 ? ba > bb : ba < bb) : cost(a) > cost(b);     // top H plane first, bottom E plane last
        });
        flat.insert(flat.end(), v.begin(), v.end());
        if (pass == 0) need[g] = (int)v.size();            // one count per finished item (group_done)
    };
    // the H groups run `lead` (+ 1 when sweeping down: E(c) also needs the group after its own) ahead of the E groups, so that
    // the E-pass items of a chunk are pulled when the H-pass items they wait for are done
    const int lead = std::max(1, env_int("SJ_TMA_LEAD", 2)) + (down ? 1 : 0);
    for (int g = 0; g < nch + lead; ++g) {
        if (g < nch) emit(0, g);
        if (g - lead >= 0) emit(1, g - lead);
    }
    CK(cudaMalloc((void **)&t.grp, need.size() * sizeof(int)));
    CK(cudaMemcpy(t.grp, need.data(), need.size() * sizeof(int), cudaMemcpyHostToDevice));     // done counters and epoch start at 0
    const int zero[4] = {0, 0, 0, 0};
    CK(cudaMalloc((void **)&t.f.items, flat.size() * sizeof(WorkItem)));
    CK(cudaMalloc((void **)&t.f.first, sizeof zero));
    CK(cudaMemcpy(t.f.items, flat.data(), flat.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t.f.first, zero, sizeof zero, cudaMemcpyHostToDevice));
    t.f.n_items = (int)flat.size(); t.f.grid = std::max(1, std::min(n_sm, (int)flat.size()));
    t.wave = wave; t.n_chunks = nch;
    return 0;
}

int sj_tma_build_materials(sj_sim *s) {
    TmaState &t = s->tma;
    if (!t.mode) return 0;
    // ---- tensor maps (F and the boxes never move; P is re-allocated when the slot count changes) ----
    std::vector<CUtensorMap> maps((size_t)t.n_shapes * SJ_TMAP_PER_SHAPE);
    memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
    const long long f_units = 6LL * s->g.n_sets * s->nzl, p_units = 2LL * std::max(s->n_slots, 1) * 3 * s->g.n_sets * s->p_nzp;
    for (int i = 0; i < t.n_shapes; ++i) {
        const TShape &sh = t.shapes[i];
        CUtensorMap *m = &maps[(size_t)i * SJ_TMAP_PER_SHAPE];
        int rc = make_map(s, m + SJ_TMAP_F_HALO, s->F, s->pitch, s->rows, f_units, s->plane, sh.hp, sh.th + 1); if (rc) return rc;
        rc = make_map(s, m + SJ_TMAP_F_OWN, s->F, s->pitch, s->rows, f_units, s->plane, sh.tw, sh.th); if (rc) return rc;
        rc = make_map(s, m + SJ_TMAP_P_OWN, s->Pall, s->pitch, s->rows, p_units, s->plane, sh.tw, sh.th); if (rc) return rc;
        for (size_t b = 0; b < s->boxes.size(); ++b) {
            const sj_sim::Box &B = s->boxes[b];
            rc = make_map(s, m + SJ_TMAP_BOX0 + b, B.base, B.bpitch, B.by, (long long)B.narr * s->g.n_sets * B.bz, B.bplane, sh.tw, sh.th); if (rc) return rc;
        }
    }
    cudaFree(t.maps); t.maps = NULL;
    CK(cudaMalloc(&t.maps, std::max<size_t>(maps.size(), 1) * sizeof(CUtensorMap)));
    CK(cudaMemcpy(t.maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    // ---- E-pass schedules: classify the geometry items by material, per tile shape ----
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    std::vector<WorkItem> cls[4], all;
    for (int sh = 0; sh < t.n_shapes; ++sh) {
        std::vector<WorkItem> sub;
        for (const WorkItem &w : t.geo[1]) if ((w.shape & 0xff) == sh) sub.push_back(w);
        int rc = sj_classify_items(s, sub, t.shapes[sh].tw, t.shapes[sh].th, cls); if (rc) return rc;
    }
    // a tall tile whose material class stages so many polarisation tiles that two plane loads would not fit the ring is
    // cut into its two half-height tiles (shape registered by sj_tma_build_geometry), and those are classified again
    {
        const int cap = env_int("SJ_TMA_RING_KB", 208) * 1024;
        std::vector<WorkItem> halves[SJ_TMA_MAX_SHAPES];
        for (int c = 1; c < 4; ++c) {
            std::vector<WorkItem> keep;
            for (const WorkItem &w : cls[c]) {
                const TShape &sh = t.shapes[w.shape & 0xff];
                const int ns = c == 1 ? std::max(s->n_slots, 1) : c - 1;
                const long long len = 3LL * sh.hs + (3 + (w.box >= 0 ? (w.kind == 0 ? 6 : 1) : 0) + 6 * ns) * (long long)sh.os;
                if (sh.sub == 1 || (ns <= 1 && 2 * len <= cap)) { keep.push_back(w); continue; }     // (the kernels take tall tiles with <= 1 slot)
                int small = -1;
                for (int i = 0; i < t.n_shapes; ++i) if (t.shapes[i].nvx == sh.nvx && t.shapes[i].th == sh.ths && t.shapes[i].sub == 1) small = i;
                if (small < 0) { s->err = "TMA schedule: half-height tile shape missing"; return SJ_ERR_ARG; }
                for (int j0 = w.j0; j0 < w.j_hi; j0 += sh.ths) {
                    WorkItem h = w; h.j0 = j0; h.j_hi = std::min(j0 + sh.ths, w.j_hi); h.shape = small | (w.shape & SJ_BND_FLAG);
                    halves[small].push_back(h);
                }
            }
            cls[c].swap(keep);
        }
        for (int sh = 0; sh < t.n_shapes; ++sh) {
            if (halves[sh].empty()) continue;
            int rc = sj_classify_items(s, halves[sh], t.shapes[sh].tw, t.shapes[sh].th, cls); if (rc) return rc;
        }
    }
    t.n_bnd[1] = 0;
    for (int c = 0; c < 4; ++c)
        for (WorkItem w : cls[c]) { w.pad = c; w.shape |= SJ_EPASS_FLAG; all.push_back(w); if (w.shape & SJ_BND_FLAG) t.n_bnd[1]++; }      // pad = material class of the item
    { int rc = upload_schedule(s, replicate_sets(all, s->g.n_sets), 1, n_sm, t.e[0][0]); if (rc) return rc; }
    return build_fused(s, all, n_sm);
}
