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
