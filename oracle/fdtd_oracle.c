/*
 * fdtd_oracle.c -- CPU restatement of the time-stepping path of sim_juncs.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file.  The product path is sim_juncs_b200/csrc (CUDA).
 *
 * What it restates
 * ----------------
 * The reference's per-step code is a single call, fields.step()
 * (reference src/disp.cpp:740), into the un-vendored third-party library
 * NanoComp/meep "v1.24" (reference README.md:8, CMakeLists.txt:24).  meep is
 * absent from /root/reference and from this image, so this file restates meep's
 * *published* algorithm (Oskooi et al., Comput. Phys. Commun. 181 (2010) 687;
 * meep src/step.cpp, step_generic.cpp, update_eh.cpp, update_pols.cpp,
 * susceptibility.cpp, structure.cpp::use_pml, sources.cpp, loop_in_chunks.cpp,
 * vec.cpp::interpolate -- recalled, not diffed) and anchors on the reference's
 * own call sites:
 *   grid                      src/disp.cpp:505-509  (meep::vol3d(L,L,L,a))
 *   eps_inf / sigma sampling  src/disp.cpp:264-316  (cgs_material_function)
 *   structure + PML           src/disp.cpp:527      (meep::pml(thickness))
 *   susceptibilities          src/disp.cpp:529-548  (lorentzian_susceptibility)
 *   source waveform           src/disp.cpp:378-400  (gaussian_src_time_phase)
 *   source placement          src/disp.cpp:603-625  (add_volume_source)
 *   run loop / sampling       src/disp.cpp:690-749  (sample Ex, then step)
 *
 * PARITY UNPINNED against real meep: the reference holds no golden field value
 *  and meep cannot be run here.  The oracle is pinned
 * instead by (i) the reference's own known-answer tests that touch this path
 * (source end time, waveform getters, config rounding, eps lookup) and (ii)
 * physics checks in tests/test_oracle_physics.py (Yee dispersion, PML
 * reflection, Lorentz/Drude slab Fresnel coefficients, energy decay).
 *
 * Conventions (meep): c = eps0 = mu0 = 1, dx = 1/a, dt = Courant/a, curl terms
 * are multiplied by Courant = dt/dx.  Arrays are (nx+1)(ny+1)(nz+1) per
 * component, x fastest.  Yee offsets: Ex(i+1/2,j,k) Ey(i,j+1/2,k) Ez(i,j,k+1/2)
 * Hx(i,j+1/2,k+1/2) Hy(i+1/2,j,k+1/2) Hz(i+1/2,j+1/2,k).  The outer wall is
 * metallic (meep default boundary): points meep does not own, or zeroes through
 * zero_metal(), are simply never updated here and stay 0.
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_SUS 32
#define ORC_MAX_SRC 8

typedef struct {
    double omega0, gamma; /* meep units (already divided by um_scale, disp.cpp:539-540) */
    int drude;            /* no_omega_0_denominator */
    int region;
    double sigma;         /* sigma/thickness, disp.cpp:541 */
    double *P[2][3], *Pp[2][3]; /* [set][comp] current and previous polarisation */
    double *sigarr[3];    /* per-point sigma (orc_add_sus_pointwise) instead of sigma * region bit, or NULL */
} orc_sus;

typedef struct {
    int comp;              /* 0..2 = Ex,Ey,Ez */
    int integrated;        /* meep src_time::is_integrated */
    int kind;              /* 0: gaussian_src_time_phase (disp.hpp:112); 1: meep::continuous_src_time (disp.cpp:618);
                              2: waveform supplied by the caller (the reference's own src_time::dipole through oracle/shim) */
    void (*fn)(void *ctx, double time, double *out2);
    void *ctx;
    double t_start, t_end, slowness;   /* kind 1 */
    double omega, width, phi, peak, cutoff;
    double _Complex amp_t; /* 1/(-i omega), disp.cpp:387 */
    double _Complex amp;   /* user amplitude times a^(zero-size dims) */
    int lo[3], hi[3];      /* index range of the component grid, inclusive */
    double *w[3];          /* per-direction integration weights over [lo,hi] */
    double _Complex cur_dipole, cur_current;
} orc_src;

typedef struct orc_sim {
    int n[3];
    size_t st[3];          /* strides */
    size_t ntot;
    double a, inva, courant, dt;
    double pml, pmlR;
    int nsets;
    long t;                /* step counter */
    /* per direction PML tables over half-pixel index 0..2n+1 */
    double *sig[3], *kap[3], *siginv[3];
    /* fields [set][comp] */
    double *E[2][3], *D[2][3], *W[2][3], *UD[2][3];
    double *H[2][3], *B[2][3], *UB[2][3], *WH[2][3];
    double *chi1inv[3];
    int have_disp;
    int n_sus;
    orc_sus sus[ORC_MAX_SUS];
    const uint8_t *mask[3]; /* borrowed during set_regions only */
    uint8_t *maskc[3];      /* owned copies */
    int n_src;
    orc_src src[ORC_MAX_SRC];
    /* monitors */
    int n_mon, mon_comp;
    double *mon_xyz;
    size_t (*mon_idx)[8];
    double (*mon_w)[8];
    int n_saves, cap_saves;
    double *series;        /* [save][mon][2] */
    int diverged;
} orc_sim;

static double *zalloc(size_t n) {
    double *p = (double *)calloc(n, sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* meep structure_chunk::use_pml: sigma(u)=prefac*u^2 sampled every half pixel,
 * prefac = -ln(R)/(4 d int_0^1 u^2 du); sig[] = sigma dt/2, kappa == 1. */
static void build_pml(orc_sim *s) {
    for (int d = 0; d < 3; ++d) {
        int m = 2 * s->n[d] + 2;
        s->sig[d] = zalloc(m); s->kap[d] = zalloc(m); s->siginv[d] = zalloc(m);
        for (int i = 0; i < m; ++i) { s->kap[d][i] = 1.0; s->siginv[d][i] = 1.0; }
        if (s->pml <= 0) continue;
        const double dx = s->pml;
        const double prefac = -log(s->pmlR) / (4 * dx * (1.0 / 3.0));
        const double bloc[2] = {0.0, (2 * s->n[d]) * (0.5 * s->inva)};
        for (int side = 0; side < 2; ++side)
            for (int i = 0; i < m; ++i) {
                double here = i * 0.5 * s->inva;
                double x = dx - (0.5 * s->inva) * ((int)(fabs(bloc[side] - here) * s->a * 2 + 0.5));
                if (x > 0) {
                    double u = x / dx;
                    s->sig[d][i] = 0.5 * s->dt * (prefac * u * u);
                    s->kap[d][i] = 1.0;
                    s->siginv[d][i] = 1 / (s->kap[d][i] + s->sig[d][i]);
                }
            }
    }
}

orc_sim *orc_create(int nx, int ny, int nz, double a, double courant, double pml_thickness,
                    double pml_R, int nsets) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    s->n[0] = nx; s->n[1] = ny; s->n[2] = nz;
    s->st[0] = 1; s->st[1] = (size_t)(nx + 1); s->st[2] = (size_t)(nx + 1) * (ny + 1);
    s->ntot = s->st[2] * (nz + 1);
    s->a = a; s->inva = 1.0 / a; s->courant = courant; s->dt = courant * s->inva;
    s->pml = pml_thickness; s->pmlR = pml_R; s->nsets = nsets < 1 ? 1 : (nsets > 2 ? 2 : nsets);
    build_pml(s);
    for (int q = 0; q < s->nsets; ++q)
        for (int c = 0; c < 3; ++c) {
            s->E[q][c] = zalloc(s->ntot); s->D[q][c] = zalloc(s->ntot);
            s->W[q][c] = zalloc(s->ntot); s->UD[q][c] = zalloc(s->ntot);
            s->H[q][c] = zalloc(s->ntot); s->B[q][c] = zalloc(s->ntot);
            s->UB[q][c] = zalloc(s->ntot); s->WH[q][c] = zalloc(s->ntot);
        }
    for (int c = 0; c < 3; ++c) {
        s->chi1inv[c] = zalloc(s->ntot);
        for (size_t i = 0; i < s->ntot; ++i) s->chi1inv[c][i] = 1.0;
    }
    return s;
}

void orc_destroy(orc_sim *s) {
    if (!s) return;
    for (int d = 0; d < 3; ++d) { free(s->sig[d]); free(s->kap[d]); free(s->siginv[d]); free(s->chi1inv[d]); free(s->maskc[d]); }
    for (int q = 0; q < 2; ++q)
        for (int c = 0; c < 3; ++c) {
            free(s->E[q][c]); free(s->D[q][c]); free(s->W[q][c]); free(s->UD[q][c]);
            free(s->H[q][c]); free(s->B[q][c]); free(s->UB[q][c]); free(s->WH[q][c]);
        }
    for (int u = 0; u < s->n_sus; ++u)
        for (int q = 0; q < 2; ++q)
            for (int c = 0; c < 3; ++c) { free(s->sus[u].P[q][c]); free(s->sus[u].Pp[q][c]); }
    for (int u = 0; u < s->n_sus; ++u) for (int c = 0; c < 3; ++c) free(s->sus[u].sigarr[c]);
    for (int i = 0; i < s->n_src; ++i) for (int d = 0; d < 3; ++d) free(s->src[i].w[d]);
    free(s->mon_xyz); free(s->mon_idx); free(s->mon_w); free(s->series);
    free(s);
}

/* Materials.  mask_c[idx] bit r = reference composite_object::in() of region r at the
 * Yee point of E component c (src/cgs.cpp:422).  eps_inf follows in_bound()
 * (src/disp.cpp:264-283) with smooth_n = 0: ret = sum_r def + (s_r - def)*in_r, i.e. the
 * ambient value is counted once per region.  Each (region,pole) is one
 * meep lorentzian_susceptibility whose sigma(r) = sigma_rp * in_r (disp.cpp:536-546).
 * poles: flat [omega0, gamma, sigma, drude] per pole, regions concatenated. */
int orc_set_regions(orc_sim *s, double ambient_eps, int n_regions, const double *region_eps,
                    const int *region_npoles, const double *poles, const uint8_t *mex,
                    const uint8_t *mey, const uint8_t *mez) {
    const uint8_t *m[3] = {mex, mey, mez};
    if (n_regions > 8) return -1;
    for (int c = 0; c < 3; ++c) {
        free(s->maskc[c]);
        s->maskc[c] = (uint8_t *)malloc(s->ntot);
        memcpy(s->maskc[c], m[c], s->ntot);
        for (size_t i = 0; i < s->ntot; ++i) {
            double ret = 0;
            if (n_regions == 0) ret = ambient_eps; /* if (!regions) return def_ret */
            for (int r = 0; r < n_regions; ++r) {
                double this_ret = (m[c][i] >> r) & 1;
                ret += ambient_eps + (region_eps[r] - ambient_eps) * this_ret / (0 + 1);
            }
            s->chi1inv[c][i] = 1 / ret; /* meep material_function::eff_chi1inv_row, maxeval = 0 */
        }
    }
    int pidx = 0;
    for (int r = 0; r < n_regions; ++r)
        for (int p = 0; p < region_npoles[r]; ++p, ++pidx) {
            if (s->n_sus >= ORC_MAX_SUS) return -2;
            orc_sus *u = &s->sus[s->n_sus++];
            u->omega0 = poles[4 * pidx + 0]; u->gamma = poles[4 * pidx + 1];
            u->sigma = poles[4 * pidx + 2]; u->drude = poles[4 * pidx + 3] != 0.0;
            u->region = r;
            for (int q = 0; q < s->nsets; ++q)
                for (int c = 0; c < 3; ++c) { u->P[q][c] = zalloc(s->ntot); u->Pp[q][c] = zalloc(s->ntot); }
            s->have_disp = 1;
        }
    return 0;
}

/* Materials with stochastic smoothing (smooth_n > 0, disp.cpp:264-283): cnt_c[r * ntot + idx] = in_r(p) + sum_j in_r(p + delta_j)
 * from csg_raster_counts, smooth_total = 8 * smooth_n offsets.  eps_inf = sum_r def + (s_r - def) * cnt / (smooth_total + 1);
 * each (region, pole) susceptibility has sigma(p) = 0 + (sigma_rp - 0) * cnt / (smooth_total + 1) (its material function is
 * built with def_ret = 0, disp.cpp:545). */
int orc_set_regions_counts(orc_sim *s, double ambient_eps, int n_regions, const double *region_eps,
                           const int *region_npoles, const double *poles, unsigned smooth_total, const uint8_t *cx,
                           const uint8_t *cy, const uint8_t *cz) {
    const uint8_t *cn[3] = {cx, cy, cz};
    for (int c = 0; c < 3; ++c) {
        free(s->maskc[c]);
        s->maskc[c] = (uint8_t *)calloc(s->ntot, 1);
        for (size_t i = 0; i < s->ntot; ++i) {
            double ret = 0;
            if (n_regions == 0) ret = ambient_eps;
            for (int r = 0; r < n_regions; ++r) {
                double this_ret = cn[c][(size_t)r * s->ntot + i];
                ret += ambient_eps + (region_eps[r] - ambient_eps) * this_ret / (smooth_total + 1);
            }
            s->chi1inv[c][i] = 1 / ret;
        }
    }
    int pidx = 0;
    for (int r = 0; r < n_regions; ++r)
        for (int p = 0; p < region_npoles[r]; ++p, ++pidx) {
            if (s->n_sus >= ORC_MAX_SUS) return -2;
            orc_sus *u = &s->sus[s->n_sus++];
            u->omega0 = poles[4 * pidx + 0]; u->gamma = poles[4 * pidx + 1];
            u->sigma = poles[4 * pidx + 2]; u->drude = poles[4 * pidx + 3] != 0.0;
            u->region = r;
            for (int c = 0; c < 3; ++c) {
                u->sigarr[c] = zalloc(s->ntot);
                for (size_t i = 0; i < s->ntot; ++i) {
                    double this_ret = cn[c][(size_t)r * s->ntot + i];
                    u->sigarr[c][i] = 0.0 + (u->sigma - 0.0) * this_ret / (smooth_total + 1);
                }
                for (int q = 0; q < s->nsets; ++q) { u->P[q][c] = zalloc(s->ntot); u->Pp[q][c] = zalloc(s->ntot); }
            }
            s->have_disp = 1;
        }
    return 0;
}

/* ---- source waveform: gaussian_src_time_phase (src/disp.cpp:378-400) ---- */
static double _Complex src_dipole(const orc_src *g, double time) {
    if (g->kind == 2) { double o[2]; g->fn(g->ctx, time, o); return o[0] + I * o[1]; }
    if (g->kind == 1) {
        /* meep continuous_src_time::dipole [meep-recall, v1.2x sources.cpp]: zero outside [start, end] (float
         * compare), exp(-i w t) / (-i w), times tanh ramps (1+tanh(ts))(1+tanh(te))/4 when width != 0 */
        float rtime = (float)time;
        if (rtime < g->t_start || rtime > g->t_end) return 0.0;
        double _Complex osc = (cos(-g->omega * time) + I * sin(-g->omega * time)) * g->amp_t;
        if (g->width == 0.0) return osc;
        double ts = (time - g->t_start) / g->width - g->slowness;
        double te = (g->t_end - time) / g->width - g->slowness;
        return osc * (1.0 + tanh(ts)) * (1.0 + tanh(te)) * 0.25;
    }
    double tt = time - g->peak;
    if ((float)fabs(tt) > g->cutoff) return 0.0;
    double _Complex pol = cos(-g->omega * tt - g->phi) + I * sin(-g->omega * tt - g->phi);
    return exp(-tt * tt / (2 * g->width * g->width)) * pol * g->amp_t;
}
/* meep src_time::update(): dipole(t) and current(t) = (dipole(t+dt)-dipole(t))/dt */
static void calc_sources(orc_sim *s, double tim) {
    for (int i = 0; i < s->n_src; ++i) {
        orc_src *g = &s->src[i];
        g->cur_dipole = src_dipole(g, tim);
        g->cur_current = (src_dipole(g, tim + s->dt) - g->cur_dipole) / s->dt;
    }
}
double orc_src_last_time(const orc_sim *s) {
    double t = 0;
    for (int i = 0; i < s->n_src; ++i) {
        double lt = s->src[i].kind == 1 ? s->src[i].t_end               /* continuous_src_time::last_time */
                                        : (float)(s->src[i].peak + s->src[i].cutoff); /* disp.hpp:119 */
        if (lt > t) t = lt;
    }
    return t;
}
void orc_src_dipole(const orc_sim *s, int isrc, double time, double *out2) {
    double _Complex d = src_dipole(&s->src[isrc], time);
    out2[0] = creal(d); out2[1] = cimag(d);
}

/* Source placement: meep fields::add_volume_source + loop_in_chunks boundary weights.
 * For each direction the volume [lo,hi] is covered by the component's grid points
 * is..ie (half-pixel indices, step 2) with linear-interpolation end weights
 * s0,s1,(1...),e1,e0; a zero-thickness direction degenerates to the two bracketing
 * planes with weights (1-f, f) and multiplies the amplitude by a (delta function). */
static int place_source(orc_sim *s, orc_src *g, int comp, const double *lo, const double *hi, double _Complex amp) {
    for (int d = 0; d < 3; ++d) {
        const int sh = (d == comp) ? 1 : 0;     /* iyee_shift of an E component */
        const int iyc = 1 - sh;                 /* iyee_shift(Centered) - iyee_shift(c) */
        const double yc = iyc * (0.5 * s->inva);
        double wmin = lo[d] + yc, wmax = hi[d] + yc;
        int is = 1 + 2 * (int)floor(wmin * s->a - .5) - iyc;
        int ie = 1 + 2 * (int)ceil(wmax * s->a - .5) - iyc;
        double w0 = 1. - lo[d] * s->a + 0.5 * is;
        double w1 = 1. + hi[d] * s->a - 0.5 * ie;
        double s0, s1, e0, e1;
        if (ie >= is + 3 * 2) {
            s0 = w0 * w0 / 2; s1 = 1 - (1 - w0) * (1 - w0) / 2;
            e0 = w1 * w1 / 2; e1 = 1 - (1 - w1) * (1 - w1) / 2;
        } else if (ie == is + 2 * 2) {
            s0 = w0 * w0 / 2; s1 = 1 - (1 - w0) * (1 - w0) / 2 - (1 - w1) * (1 - w1) / 2;
            e0 = w1 * w1 / 2; e1 = s1;
        } else if (lo[d] == hi[d]) {
            s0 = w0; s1 = w1; e0 = w1; e1 = w0;
        } else if (ie == is + 1 * 2) {
            s0 = w0 * w0 / 2 - (1 - w1) * (1 - w1) / 2;
            e0 = w1 * w1 / 2 - (1 - w0) * (1 - w0) / 2;
            s1 = e0; e1 = s0;
        } else {
            return -2;
        }
        if (lo[d] == hi[d]) amp *= s->a;
        /* owned index range of this component in direction d */
        const int own_lo = sh ? 0 : 1, own_hi = sh ? s->n[d] - 1 : s->n[d] - 1; /* index n is metal */
        int i0 = (is - sh) / 2, i1 = (ie - sh) / 2; /* exact: is,ie have parity sh */
        int a0 = i0 < own_lo ? own_lo : i0, a1 = i1 > own_hi ? own_hi : i1;
        g->lo[d] = a0; g->hi[d] = a1;
        int cnt = a1 >= a0 ? a1 - a0 + 1 : 0;
        g->w[d] = zalloc(cnt > 0 ? cnt : 1);
        for (int i = a0; i <= a1; ++i) {
            int h = 2 * i + sh;
            double w = (h == is) ? s0 : (h == is + 2) ? s1 : (h == ie) ? e0 : (h == ie - 2) ? e1 : 1.0;
            g->w[d][i - a0] = w;
        }
    }
    g->amp = amp;
    s->n_src++;
    return 0;
}

int orc_add_gaussian_source(orc_sim *s, int comp, const double *lo, const double *hi,
                            double amp_re, double amp_im, double freq, double width, double phase,
                            double t_start, double t_end, int integrated) {
    if (s->n_src >= ORC_MAX_SRC || comp < 0 || comp > 2) return -1;
    orc_src *g = &s->src[s->n_src];
    memset(g, 0, sizeof(*g));
    g->comp = comp; g->integrated = integrated; g->kind = 0;
    g->omega = 2 * M_PI * freq; g->width = width; g->phi = phase + M_PI;
    g->peak = 0.5 * (t_start + t_end); g->cutoff = (t_end - t_start) * 0.5;
    g->amp_t = 1.0 / (0.0 - I * g->omega);
    while (exp(-g->cutoff * g->cutoff / (2 * g->width * g->width)) < 1e-100) g->cutoff *= 0.9;
    g->cutoff = (float)g->cutoff;
    return place_source(s, g, comp, lo, hi, amp_re + I * amp_im);
}

/* meep::continuous_src_time(freq, width, start, end, slowness) as the reference builds it (disp.cpp:618; the
 * reference leaves slowness at meep's default 3.0 and feeds the scene's `slowness` argument in as the width) */
int orc_add_cw_source(orc_sim *s, int comp, const double *lo, const double *hi, double amp_re, double amp_im,
                      double freq, double width, double t_start, double t_end, double slowness, int integrated) {
    if (s->n_src >= ORC_MAX_SRC || comp < 0 || comp > 2) return -1;
    orc_src *g = &s->src[s->n_src];
    memset(g, 0, sizeof(*g));
    g->comp = comp; g->integrated = integrated; g->kind = 1;
    g->omega = 2 * M_PI * freq; g->width = width; g->t_start = t_start; g->t_end = t_end; g->slowness = slowness;
    g->amp_t = 1.0 / (0.0 - I * g->omega);
    return place_source(s, g, comp, lo, hi, amp_re + I * amp_im);
}

/* Monitors: meep fields::get_field -> grid_volume::interpolate (linear in each direction
 * between the two bracketing Yee points of the component). */
/* ---- entry points for the meep-API shim (oracle/shim/): eps_inf, sigma and the source waveform come from the
 * caller point by point -- there they are the reference's own cgs_material_function / src_time virtuals ---- */
int orc_set_eps_pointwise(orc_sim *s, const double *ex, const double *ey, const double *ez) {
    const double *e[3] = {ex, ey, ez};
    for (int c = 0; c < 3; ++c) {
        free(s->maskc[c]);
        s->maskc[c] = (uint8_t *)calloc(s->ntot, 1);
        for (size_t i = 0; i < s->ntot; ++i) s->chi1inv[c][i] = 1 / e[c][i];
    }
    return 0;
}
int orc_add_sus_pointwise(orc_sim *s, double omega0, double gamma, int drude, const double *sx, const double *sy,
                          const double *sz) {
    const double *sg[3] = {sx, sy, sz};
    if (s->n_sus >= ORC_MAX_SUS) return -2;
    orc_sus *u = &s->sus[s->n_sus++];
    u->omega0 = omega0; u->gamma = gamma; u->sigma = 0; u->drude = drude; u->region = 0;
    for (int c = 0; c < 3; ++c) {
        u->sigarr[c] = zalloc(s->ntot);
        memcpy(u->sigarr[c], sg[c], sizeof(double) * s->ntot);
        for (int q = 0; q < s->nsets; ++q) { u->P[q][c] = zalloc(s->ntot); u->Pp[q][c] = zalloc(s->ntot); }
    }
    s->have_disp = 1;
    return 0;
}
int orc_add_callback_source(orc_sim *s, int comp, const double *lo, const double *hi, double amp_re, double amp_im,
                            int integrated, void (*fn)(void *, double, double *), void *ctx) {
    if (s->n_src >= ORC_MAX_SRC || comp < 0 || comp > 2) return -1;
    orc_src *g = &s->src[s->n_src];
    memset(g, 0, sizeof(*g));
    g->comp = comp; g->integrated = integrated; g->kind = 2; g->fn = fn; g->ctx = ctx;
    return place_source(s, g, comp, lo, hi, amp_re + I * amp_im);
}

/* meep fields::get_field(c, loc): (tri)linear interpolation from the <= 8 surrounding points of component c */
static void interp_weights(const orc_sim *s, int comp, const double *xyz, size_t *idx8, double *w8) {
    int mid[3]; double dv[3];
    for (int d = 0; d < 3; ++d) {
        int sh = (d == comp) ? 1 : 0;
        double pc = xyz[d];
        double p = (pc - sh * (0.5 * s->inva)) * s->a;
        mid[d] = ((int)floor(p)) * 2 + 1 + sh;
        double midv = mid[d] * (0.5 * s->inva);
        dv[d] = (pc - midv) * (2 * s->a);
    }
    for (int q = 0; q < 8; ++q) {
        double w = 1.0; size_t idx = 0; int ok = 1;
        for (int d = 0; d < 3; ++d) {
            int sh = (d == comp) ? 1 : 0;
            int hi = (q >> d) & 1;
            int h = mid[d] + (hi ? 1 : -1);
            w *= hi ? 0.5 * (1.0 + dv[d]) : 0.5 * (1.0 - dv[d]);
            int i = (h - sh) / 2;
            if (h - sh < 0 || i > s->n[d]) ok = 0; else idx += (size_t)i * s->st[d];
        }
        if (w < 0.0) w = 0.0;
        if (!ok) { w = 0.0; idx = 0; }
        idx8[q] = idx; w8[q] = w;
    }
}

int orc_add_monitors(orc_sim *s, int comp, int n, const double *xyz) {
    s->mon_comp = comp; s->n_mon = n;
    s->mon_xyz = (double *)malloc(sizeof(double) * 3 * n);
    memcpy(s->mon_xyz, xyz, sizeof(double) * 3 * n);
    s->mon_idx = malloc(sizeof(size_t[8]) * n);
    s->mon_w = malloc(sizeof(double[8]) * n);
    for (int m = 0; m < n; ++m) interp_weights(s, comp, xyz + 3 * m, s->mon_idx[m], s->mon_w[m]);
    return 0;
}

/* one get_field call, same summation order as sample_monitors */
void orc_get_field_at(const orc_sim *s, int comp, const double *xyz, double *out2) {
    size_t idx[8]; double w[8];
    interp_weights(s, comp, xyz, idx, w);
    for (int q = 0; q < 2; ++q) {
        double res = 0.0;
        if (q < s->nsets)
            for (int p = 0; p < 8; ++p)
                if (w[p] != 0) res += w[p] * s->E[q][comp][idx[p]];
        out2[q] = res;
    }
}

static void sample_monitors(orc_sim *s) {
    if (s->n_mon == 0) return;
    if (s->n_saves >= s->cap_saves) {
        s->cap_saves = s->cap_saves ? 2 * s->cap_saves : 256;
        s->series = (double *)realloc(s->series, sizeof(double) * 2 * s->n_mon * s->cap_saves);
    }
    double *row = s->series + (size_t)2 * s->n_mon * s->n_saves;
    for (int m = 0; m < s->n_mon; ++m)
        for (int q = 0; q < 2; ++q) {
            double res = 0.0;
            if (q < s->nsets)
                for (int p = 0; p < 8; ++p)
                    if (s->mon_w[m][p] != 0) res += s->mon_w[m][p] * s->E[q][s->mon_comp][s->mon_idx[m][p]];
            row[2 * m + q] = res;
            if (q == 0 && res > 1000) s->diverged = 1; /* disp.cpp:727 */
        }
    s->n_saves++;
}

/* meep step_curl, "fu update + PML in f update" branch with kappa tables; where a sigma is
 * zero the expression degenerates exactly to the simpler branches. */
static void step_db(orc_sim *s, int is_D) {
    const double C = s->courant;
    for (int q = 0; q < s->nsets; ++q)
        for (int c = 0; c < 3; ++c) {
            const int d1 = (c + 1) % 3, d2 = (c + 2) % 3;
            double *f = is_D ? s->D[q][c] : s->B[q][c];
            double *fu = is_D ? s->UD[q][c] : s->UB[q][c];
            const double *g1 = is_D ? s->H[q][d2] : s->E[q][d2]; /* derivative along d1 */
            const double *g2 = is_D ? s->H[q][d1] : s->E[q][d1]; /* derivative along d2 */
            const ptrdiff_t s1 = (is_D ? -1 : 1) * (ptrdiff_t)s->st[d1];
            const ptrdiff_t s2 = (is_D ? -1 : 1) * (ptrdiff_t)s->st[d2];
            int lo[3], hi[3], sh[3];
            for (int d = 0; d < 3; ++d) {
                if (is_D) { sh[d] = (d == c); lo[d] = sh[d] ? 0 : 1; hi[d] = s->n[d] - 1; }
                else      { sh[d] = (d != c); lo[d] = sh[d] ? 0 : 1; hi[d] = s->n[d] - 1; }
            }
            const double *sigk = s->sig[d1], *kapk = s->kap[d1], *sinvk = s->siginv[d1];
            const double *sigu = s->sig[d2], *kapu = s->kap[d2], *sinvu = s->siginv[d2];
#pragma omp parallel for collapse(2) schedule(static)
            for (int k = lo[2]; k <= hi[2]; ++k)
                for (int j = lo[1]; j <= hi[1]; ++j) {
                    int ijk[3]; ijk[2] = k; ijk[1] = j;
                    for (int i = lo[0]; i <= hi[0]; ++i) {
                        ijk[0] = i;
                        size_t x = (size_t)k * s->st[2] + (size_t)j * s->st[1] + i;
                        int hk = 2 * ijk[d1] + sh[d1], hu = 2 * ijk[d2] + sh[d2];
                        double curl = C * (g1[x + s1] - g1[x] + g2[x] - g2[x + s2]);
                        if (sigu[hu] != 0.0) {
                            double fprev = fu[x];
                            fu[x] = ((kapk[hk] - sigk[hk]) * fu[x] - curl) * sinvk[hk];
                            f[x] = ((kapu[hu] - sigu[hu]) * f[x] + (fu[x] - fprev)) * sinvu[hu];
                        } else {
                            f[x] = ((kapk[hk] - sigk[hk]) * f[x] - curl) * sinvk[hk];
                        }
                    }
                }
        }
}

/* meep fields_chunk::step_source for non-integrated sources: D -= dt * A * current */
static void step_source_D(orc_sim *s) {
    for (int n = 0; n < s->n_src; ++n) {
        orc_src *g = &s->src[n];
        if (g->integrated) continue;
        for (int k = g->lo[2]; k <= g->hi[2]; ++k)
            for (int j = g->lo[1]; j <= g->hi[1]; ++j)
                for (int i = g->lo[0]; i <= g->hi[0]; ++i) {
                    size_t x = (size_t)k * s->st[2] + (size_t)j * s->st[1] + i;
                    double wgt = g->w[0][i - g->lo[0]] * g->w[1][j - g->lo[1]] * g->w[2][k - g->lo[2]];
                    double _Complex A = (wgt * g->amp) * g->cur_current * s->dt;
                    s->D[0][g->comp][x] -= creal(A);
                    if (s->nsets > 1) s->D[1][g->comp][x] -= cimag(A);
                }
    }
}

/* meep update_eh(H_stuff): mu == 1, so the auxiliary W_H is a copy of B and
 * H += (kap+sig) W_new - (kap-sig) W_old along the component's own direction (dsigw = d_c). */
static void update_h(orc_sim *s) {
    for (int q = 0; q < s->nsets; ++q)
        for (int c = 0; c < 3; ++c) {
            const double *sigw = s->sig[c], *kapw = s->kap[c];
            int lo[3], hi[3], sh[3];
            for (int d = 0; d < 3; ++d) { sh[d] = (d != c); lo[d] = sh[d] ? 0 : 1; hi[d] = s->n[d] - 1; }
            double *H = s->H[q][c], *W = s->WH[q][c]; const double *B = s->B[q][c];
#pragma omp parallel for collapse(2) schedule(static)
            for (int k = lo[2]; k <= hi[2]; ++k)
                for (int j = lo[1]; j <= hi[1]; ++j) {
                    int ijk[3]; ijk[2] = k; ijk[1] = j;
                    for (int i = lo[0]; i <= hi[0]; ++i) {
                        ijk[0] = i;
                        size_t x = (size_t)k * s->st[2] + (size_t)j * s->st[1] + i;
                        int hw = 2 * ijk[c] + sh[c];
                        double fwprev = W[x];
                        W[x] = B[x];
                        if (sigw[hw] != 0.0)
                            H[x] += (kapw[hw] + sigw[hw]) * W[x] - (kapw[hw] - sigw[hw]) * fwprev;
                        else
                            H[x] = W[x];
                    }
                }
        }
}

/* meep update_eh(E_stuff): f_minus_p = D - sum P - integrated-source dipoles;
 * W = chi1inv * f_minus_p; E += (kap+sig) W - (kap-sig) W_old (or E = W outside PML).
 * Then update_pols(E_stuff): lorentzian_susceptibility::update_P driven by W. */
static void update_e(orc_sim *s) {
    for (int q = 0; q < s->nsets; ++q)
        for (int c = 0; c < 3; ++c) {
            const double *sigw = s->sig[c], *kapw = s->kap[c];
            int lo[3], hi[3], sh[3];
            for (int d = 0; d < 3; ++d) { sh[d] = (d == c); lo[d] = sh[d] ? 0 : 1; hi[d] = s->n[d] - 1; }
            double *E = s->E[q][c], *W = s->W[q][c]; const double *D = s->D[q][c];
            const double *u = s->chi1inv[c];
#pragma omp parallel for collapse(2) schedule(static)
            for (int k = lo[2]; k <= hi[2]; ++k)
                for (int j = lo[1]; j <= hi[1]; ++j) {
                    int ijk[3]; ijk[2] = k; ijk[1] = j;
                    for (int i = lo[0]; i <= hi[0]; ++i) {
                        ijk[0] = i;
                        size_t x = (size_t)k * s->st[2] + (size_t)j * s->st[1] + i;
                        double fmp = D[x];
                        for (int n = 0; n < s->n_sus; ++n) fmp -= s->sus[n].P[q][c][x];
                        for (int n = 0; n < s->n_src; ++n) {
                            const orc_src *g = &s->src[n];
                            if (!g->integrated || g->comp != c) continue;
                            if (i < g->lo[0] || i > g->hi[0] || j < g->lo[1] || j > g->hi[1] || k < g->lo[2] || k > g->hi[2]) continue;
                            double wgt = g->w[0][i - g->lo[0]] * g->w[1][j - g->lo[1]] * g->w[2][k - g->lo[2]];
                            double _Complex A = (wgt * g->amp) * g->cur_dipole;
                            fmp -= q ? cimag(A) : creal(A);
                        }
                        int hw = 2 * ijk[c] + sh[c];
                        double wnew = fmp * u[x];
                        if (sigw[hw] != 0.0) {
                            double wprev = W[x];
                            W[x] = wnew;
                            E[x] += (kapw[hw] + sigw[hw]) * wnew - (kapw[hw] - sigw[hw]) * wprev;
                        } else {
                            W[x] = wnew;
                            E[x] = wnew;
                        }
                    }
                }
        }
    /* update_pols */
    const double dt = s->dt;
    for (int n = 0; n < s->n_sus; ++n) {
        orc_sus *su = &s->sus[n];
        const double omega2pi = 2 * M_PI * su->omega0, g2pi = su->gamma * 2 * M_PI;
        const double omega0dtsqr = omega2pi * omega2pi * dt * dt;
        const double gamma1inv = 1 / (1 + g2pi * dt / 2), gamma1 = (1 - g2pi * dt / 2);
        const double omega0dtsqr_denom = su->drude ? 0 : omega0dtsqr;
        for (int q = 0; q < s->nsets; ++q)
            for (int c = 0; c < 3; ++c) {
                int lo[3], hi[3];
                for (int d = 0; d < 3; ++d) { int sh = (d == c); lo[d] = sh ? 0 : 1; hi[d] = s->n[d] - 1; }
                double *p = su->P[q][c], *pp = su->Pp[q][c];
                const double *w = s->W[q][c];
                const uint8_t *mk = s->maskc[c];
#pragma omp parallel for collapse(2) schedule(static)
                for (int k = lo[2]; k <= hi[2]; ++k)
                    for (int j = lo[1]; j <= hi[1]; ++j)
                        for (int i = lo[0]; i <= hi[0]; ++i) {
                            size_t x = (size_t)k * s->st[2] + (size_t)j * s->st[1] + i;
                            double sg = su->sigarr[c] ? su->sigarr[c][x] : su->sigma * ((mk[x] >> su->region) & 1);
                            double pcur = p[x];
                            p[x] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[x] + omega0dtsqr * (sg * w[x]));
                            pp[x] = pcur;
                        }
            }
    }
}

/* One meep fields::step() (src/disp.cpp:740). */
void orc_step(orc_sim *s) {
    const double time = s->t * s->dt;
    calc_sources(s, time);
    step_db(s, 0);
    calc_sources(s, time + 0.5 * s->dt);
    update_h(s);
    calc_sources(s, time + 0.5 * s->dt);
    step_db(s, 1);
    step_source_D(s);
    calc_sources(s, time + s->dt);
    update_e(s);
    s->t += 1;
}

/* The reference run loop (src/disp.cpp:719-741): sample, then step. */
int orc_run(orc_sim *s, int n_steps, int save_span) {
    if (save_span == 0) save_span = 1;
    for (int i = 0; i < n_steps; ++i) {
        if (i % save_span == 0) sample_monitors(s);
        orc_step(s);
    }
    return s->diverged;
}

int orc_n_saves(const orc_sim *s) { return s->n_saves; }
void orc_read_monitors(const orc_sim *s, double *out) {
    memcpy(out, s->series, sizeof(double) * 2 * s->n_mon * s->n_saves);
}
/* kind: 0 E, 1 H, 2 D, 3 B, 4 W, 5 chi1inv */
double *orc_field(orc_sim *s, int kind, int comp, int set) {
    switch (kind) {
        case 0: return s->E[set][comp];
        case 1: return s->H[set][comp];
        case 2: return s->D[set][comp];
        case 3: return s->B[set][comp];
        case 4: return s->W[set][comp];
        case 5: return s->chi1inv[comp];
    }
    return NULL;
}
double *orc_pol(orc_sim *s, int isus, int comp, int set) { return s->sus[isus].P[set][comp]; }
double orc_dt(const orc_sim *s) { return s->dt; }
long orc_time_steps(const orc_sim *s) { return s->t; }
const double *orc_pml_sig(const orc_sim *s, int d) { return s->sig[d]; }
int orc_src_range(const orc_sim *s, int isrc, int *lo, int *hi) {
    for (int d = 0; d < 3; ++d) { lo[d] = s->src[isrc].lo[d]; hi[d] = s->src[isrc].hi[d]; }
    return 0;
}
const double *orc_src_weights(const orc_sim *s, int isrc, int d) { return s->src[isrc].w[d]; }
void orc_src_amp(const orc_sim *s, int isrc, double *out2) { out2[0] = creal(s->src[isrc].amp); out2[1] = cimag(s->src[isrc].amp); }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
