/*
 * csg_oracle.c -- CPU restatement of the reference's CSG inside tests (voxelization half of
 * the hot path).  TEST INFRASTRUCTURE ONLY (see fdtd_oracle.c header).
 *
 * Follows, operation for operation and in the same floating-point order:
 *   matrix<3,3>*matrix<3,1>         reference src/geometry.hpp:88-100  (ret=0; ret+=m[i][k]*v[k])
 *   vec3::dot, rvector::normsq/norm reference src/geometry.hpp:164-173,199-201
 *   sphere::in                      reference src/cgs.cpp:34-39
 *   box::in                         reference src/cgs.cpp:52-67
 *   plane::in                       reference src/cgs.cpp:92-98
 *   cylinder::in                    reference src/cgs.cpp:110-118
 *   composite_object::in / call_child_in   reference src/cgs.cpp:403-447
 * Pinned by: the six golden inside tests of reference src/main_test.cpp:1556-1561, the eps
 * lookups of src/main_test.cpp:1635-1640, and bit-for-bit comparison with the compiled reference
 * (oracle/_ref/scene_dump points) over whole Yee grids (tests/golden/*.npz).
 * Build with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

enum { CSG_COMPOSITE = 0, CSG_SPHERE = 1, CSG_BOX = 2, CSG_PLANE = 3, CSG_CYLINDER = 4, CSG_UNDEF = 5 };
enum { CMB_UNION = 0, CMB_INTERSECT = 1, CMB_DIFFERENCE = 2, CMB_NOOP = 3 };

typedef struct {
    int32_t type, invert, cmb, child0, child1, pad;
    double M[9];
    /* sphere: center[3], rad | box: center[3], offset[3] | plane: normal[3], offset |
     * cylinder: center[3], height, r1_sq, r1_sq_x_h, r2_sq */
    double p[7];
} csg_node;

static void matvec(const double *M, const double *v, double *out) {
    for (int i = 0; i < 3; ++i) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += M[3 * i + k] * v[k];
        out[i] = acc;
    }
}

static int node_in(const csg_node *nodes, int idx, const double *r, int parent_invert) {
    const csg_node *n = &nodes[idx];
    double d[3], rel[3];
    switch (n->type) {
        case CSG_SPHERE: {
            for (int i = 0; i < 3; ++i) d[i] = r[i] - n->p[i];
            matvec(n->M, d, rel);
            double nsq = 0.0;
            for (int i = 0; i < 3; ++i) nsq += rel[i] * rel[i];
            if (sqrt(nsq) > n->p[3]) return n->invert;
            return 1 - n->invert;
        }
        case CSG_BOX: {
            for (int i = 0; i < 3; ++i) d[i] = r[i] - n->p[i];
            matvec(n->M, d, rel);
            if (fabs(rel[0]) > n->p[3]) return n->invert;
            if (fabs(rel[1]) > n->p[4]) return n->invert;
            if (fabs(rel[2]) > n->p[5]) return n->invert;
            return 1 - n->invert;
        }
        case CSG_PLANE: {
            double norm_comp = r[0] * n->p[0] + r[1] * n->p[1] + r[2] * n->p[2];
            return norm_comp > n->p[3] ? 0 : 1;
        }
        case CSG_CYLINDER: {
            for (int i = 0; i < 3; ++i) d[i] = r[i] - n->p[i];
            matvec(n->M, d, rel);
            const double height = n->p[3], r1_sq = n->p[4], r1_sq_x_h = n->p[5], r2_sq = n->p[6];
            if (rel[2] < 0 || rel[2] > height) return n->invert;
            double cent_rad_sq = rel[0] * rel[0] + rel[1] * rel[1];
            if (cent_rad_sq * height > (r2_sq - r1_sq) * rel[2] + r1_sq_x_h) return n->invert;
            return 1 - n->invert;
        }
        case CSG_COMPOSITE: {
            double rt[3];
            matvec(n->M, r, rt);
            if (n->cmb == CMB_NOOP) return 0;
            if (n->child1 < 0) {
                if (n->child0 < 0) return n->invert;
                return n->invert ^ node_in(nodes, n->child0, rt, n->invert);
            }
            if (n->child0 < 0) return n->invert ^ node_in(nodes, n->child1, rt, n->invert);
            int l = node_in(nodes, n->child0, rt, n->invert);
            int rr = node_in(nodes, n->child1, rt, n->invert);
            if (n->cmb == CMB_UNION) return n->invert ^ (l | rr);
            if (n->cmb == CMB_INTERSECT || n->cmb == CMB_DIFFERENCE) return n->invert ^ (l & rr);
            return n->invert;
        }
        default: /* call_child_in falls through to `return invert` of the calling composite */
            return parent_invert;
    }
}

/* out[p] bit r = roots[r] contains point p */
void csg_eval_points(const csg_node *nodes, const int32_t *roots, int n_roots, const double *xyz,
                     size_t npts, uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t p = 0; p < (ptrdiff_t)npts; ++p) {
        uint8_t m = 0;
        for (int r = 0; r < n_roots && r < 8; ++r)
            if (node_in(nodes, roots[r], xyz + 3 * p, 0)) m |= (uint8_t)(1u << r);
        out[p] = m;
    }
}

/* Yee sample coordinates as meep forms them (recalled): h = m*(0.5*inva) for half-pixel index m
 * (grid_volume::operator[]); the eps sample is the centre of the pixel volume dV around it,
 * ((h-q)+(h+q))*0.5 with q = 0.5*inva (structure_chunk::set_chi1inv -> gv.dV(here).center());
 * the susceptibility sigma sample is IVEC_LOOP_LOC = (0.5*is+i)*inva == h exactly. */
double csg_yee_coord(int m, double a, int kind) {
    const double inva = 1.0 / a;
    const double h = m * (0.5 * inva);
    if (kind == 0) return h;
    const double q = 0.5 * inva * 1.0;
    return ((h - q) + (h + q)) * 0.5;
}

/* Rasterize one E component (comp 0..2) of an (nx+1)(ny+1)(nz+1) Yee grid: out bit r. */
void csg_raster_component(const csg_node *nodes, const int32_t *roots, int n_roots, int nx, int ny,
                          int nz, double a, int comp, int coord_kind, uint8_t *out) {
    const int n[3] = {nx, ny, nz};
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= n[2]; ++k)
        for (int j = 0; j <= n[1]; ++j)
            for (int i = 0; i <= n[0]; ++i) {
                double r[3];
                r[0] = csg_yee_coord(2 * i + (comp == 0), a, coord_kind);
                r[1] = csg_yee_coord(2 * j + (comp == 1), a, coord_kind);
                r[2] = csg_yee_coord(2 * k + (comp == 2), a, coord_kind);
                uint8_t m = 0;
                for (int q = 0; q < n_roots && q < 8; ++q)
                    if (node_in(nodes, roots[q], r, 0)) m |= (uint8_t)(1u << q);
                out[((size_t)k * sy + j) * sx + i] = m;
            }
}

/* ---------------------------------------------------------------------------------------------------
 * Stochastic boundary smoothing (reference src/disp.cpp:56-112, generate_smooth_pts) -- restated.
 * The reference draws smooth_n offsets with libstdc++'s <random>:
 *     std::seed_seq{seed & 0xffff, (seed & 0xffff0000) >> 32}   with seed = 0xd9a28bf3 -> {0x8bf3, 0}
 *     std::mt19937 gen(seeder); std::normal_distribution<double>(0, rad); std::uniform_real_distribution<double>(0, 1)
 * per point:  theta_inv = unif(gen); phi = 2 pi unif(gen); r = gaussian(gen);
 *             x = r sin_theta cos(phi), y = r sin_theta cos(phi) (sic: cos, like x), z = r cos_theta
 * and stores the 8 sign reflections, x sign slowest.  Below: the C++11 standard's seed_seq::generate and mt19937,
 * and libstdc++'s generate_canonical<double, 53> (two 32-bit draws, low word first) and normal_distribution
 * (Marsaglia polar method, the second variate saved for the next call) -- the published algorithms, not library code.
 * Pinned by tests/test_oracle_csg.py against a program that calls <random> itself and against eps fields produced by
 * the reference's own in_bound (tests/golden/ref_run_slabs_smooth*.npz).
 * ------------------------------------------------------------------------------------------------- */
typedef struct { uint32_t x[624]; int p; int saved_ok; double saved; } smooth_rng;

static void rng_seed_seq(smooth_rng *g, const uint32_t *v, int s) {
    enum { n = 624 };
    uint32_t *b = g->x;
    for (int i = 0; i < n; ++i) b[i] = 0x8b8b8b8bu;
    const int t = 11, p = (n - t) / 2, q = p + t;
    const int m = (s + 1 > n) ? s + 1 : n;
    for (int k = 0; k < m; ++k) {
        uint32_t a = b[k % n] ^ b[(k + p) % n] ^ b[(k + n - 1) % n];
        uint32_t r1 = 1664525u * (a ^ (a >> 27));
        uint32_t r2 = r1 + (k == 0 ? (uint32_t)s : (k <= s ? (uint32_t)(k % n) + v[k - 1] : (uint32_t)(k % n)));
        b[(k + p) % n] += r1; b[(k + q) % n] += r2; b[k % n] = r2;
    }
    for (int k = m; k < m + n; ++k) {
        uint32_t a = b[k % n] + b[(k + p) % n] + b[(k + n - 1) % n];
        uint32_t r3 = 1566083941u * (a ^ (a >> 27));
        uint32_t r4 = r3 - (uint32_t)(k % n);
        b[(k + p) % n] ^= r3; b[(k + q) % n] ^= r4; b[k % n] = r4;
    }
    /* mersenne_twister_engine::seed(Sseq&): an all-zero state (top bit of x[0], every other word) becomes 2^31 */
    int zero = (b[0] & 0x80000000u) == 0;
    for (int i = 1; i < n && zero; ++i) if (b[i] != 0) zero = 0;
    if (zero) b[0] = 0x80000000u;
    g->p = n; g->saved_ok = 0; g->saved = 0;
}
static uint32_t rng_next(smooth_rng *g) {
    enum { n = 624, m = 397 };
    if (g->p >= n) {
        uint32_t *x = g->x;
        for (int k = 0; k < n; ++k) {
            uint32_t y = (x[k] & 0x80000000u) | (x[(k + 1) % n] & 0x7fffffffu);
            x[k] = x[(k + m) % n] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->p = 0;
    }
    uint32_t z = g->x[g->p++];
    z ^= (z >> 11);
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= (z >> 18);
    return z;
}
static double rng_canonical(smooth_rng *g) {      /* generate_canonical<double, 53>: k = 2 draws of 32 bits */
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) { sum += (double)rng_next(g) * tmp; tmp *= 4294967296.0; }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}
static double rng_normal(smooth_rng *g, double mean, double stddev) {
    double ret;
    if (g->saved_ok) { g->saved_ok = 0; ret = g->saved; }
    else {
        double x, y, r2;
        do {
            x = 2.0 * rng_canonical(g) - 1.0;
            y = 2.0 * rng_canonical(g) - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = sqrt(-2 * log(r2) / r2);
        g->saved = x * mult; g->saved_ok = 1;
        ret = y * mult;
    }
    return ret * stddev + mean;
}

/* out: 24 * smooth_n doubles = 8 * smooth_n offset vectors; returns 8 * smooth_n (the reference's final smooth_n) */
int csg_smooth_points(int smooth_n, double smooth_rad, double *out) {
    const uint32_t seed_words[2] = {0x8bf3u, 0u};        /* DEF_SEED = 0xd9a28bf3 through disp.cpp:58 */
    smooth_rng g;
    rng_seed_seq(&g, seed_words, 2);
    for (int i = 0; i < smooth_n; ++i) {
        int j = 0;
        double theta_inv = rng_canonical(&g) * (1.0 - 0.0) + 0.0;
        double cos_theta = 1 - 2 * theta_inv;
        double sin_theta = 2 * sqrt(theta_inv * (1 - theta_inv));
        double phi = 2 * M_PI * (rng_canonical(&g) * (1.0 - 0.0) + 0.0);
        double r = rng_normal(&g, 0.0, smooth_rad);
        double x = r * sin_theta * cos(phi);
        double y = r * sin_theta * cos(phi);             /* sic, disp.cpp:92 */
        double z = r * cos_theta;
        for (int xf = -1; xf < 2; xf += 2)
            for (int yf = -1; yf < 2; yf += 2)
                for (int zf = -1; zf < 2; zf += 2) {
                    out[24 * i + 3 * j] = x * xf; out[24 * i + 3 * j + 1] = y * yf; out[24 * i + 3 * j + 2] = z * zf;
                    ++j;
                }
    }
    return 8 * smooth_n;
}

/* in_bound's inner sums (disp.cpp:271-279): counts[r][point] = in_r(p) + sum_j in_r(p + delta_j), at the Yee points
 * of E component comp.  counts: n_roots arrays of (nx+1)(ny+1)(nz+1) bytes, region-major. */
void csg_raster_counts(const csg_node *nodes, const int32_t *roots, int n_roots, int nx, int ny, int nz, double a,
                       int comp, const double *pts, int n_pts, uint8_t *counts) {
    const size_t sx = (size_t)nx + 1, sy = (size_t)ny + 1, ntot = sx * sy * ((size_t)nz + 1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) {
                const double r_x = csg_yee_coord(2 * i + (comp == 0), a, 0);
                const double r_y = csg_yee_coord(2 * j + (comp == 1), a, 0);
                const double r_z = csg_yee_coord(2 * k + (comp == 2), a, 0);
                for (int q = 0; q < n_roots; ++q) {
                    double r[3] = {r_x, r_y, r_z};
                    int c = node_in(nodes, roots[q], r, 0);
                    for (int s = 0; s < n_pts; ++s) {
                        double rs[3] = {r_x + pts[3 * s], r_y + pts[3 * s + 1], r_z + pts[3 * s + 2]};
                        c += node_in(nodes, roots[q], rs, 0);
                    }
                    counts[(size_t)q * ntot + ((size_t)k * sy + j) * sx + i] = (uint8_t)c;
                }
            }
}
