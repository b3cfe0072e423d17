// sj_raster.cu -- one-time GPU rasterization of the CSG scene into per-component region masks
// (SURVEY.md section 8a rows A5/A6).  Replaces the >= 9 N^3 virtual cgs_material_function::in_bound
// calls meep makes through chi1p1()/sigma_row() (reference src/disp.cpp:264-306), each of which
// walks composite_object::in (src/cgs.cpp:422-447).
//
// Bit-exactness contract: every floating-point operation is performed in fp64 in the reference's
// order with NO fused multiply-add -- this translation unit is compiled with -fmad=false and only
// uses IEEE +,-,*,sqrt.  The reference is built for baseline x86-64 (SSE2, no FMA).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <random>
#include <vector>

#include "sj_internal.h"

int sj_finish_materials(sj_sim *s);

#define RCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            s->err = b_;                                                                           \
            return SJ_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define SJ_MAX_NODES 256
#define SJ_MAX_DEPTH 24
__constant__ sj_csg_node c_nodes[SJ_MAX_NODES];
__constant__ int c_roots[8];

__device__ __forceinline__ void matvec(const double *M, const double *v, double *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc += M[3 * i + k] * v[k];   // geometry.hpp:88-100
        o[i] = acc;
    }
}

__device__ int prim_in(const sj_csg_node &n, const double *r) {
    double d[3], rel[3];
    switch (n.type) {
        case 1: {  // sphere::in, cgs.cpp:34-39
            for (int i = 0; i < 3; ++i) d[i] = r[i] - n.p[i];
            matvec(n.M, d, rel);
            double nsq = 0.0;
            for (int i = 0; i < 3; ++i) nsq += rel[i] * rel[i];
            return (sqrt(nsq) > n.p[3]) ? n.invert : 1 - n.invert;
        }
        case 2: {  // box::in, cgs.cpp:52-67
            for (int i = 0; i < 3; ++i) d[i] = r[i] - n.p[i];
            matvec(n.M, d, rel);
            if (fabs(rel[0]) > n.p[3]) return n.invert;
            if (fabs(rel[1]) > n.p[4]) return n.invert;
            if (fabs(rel[2]) > n.p[5]) return n.invert;
            return 1 - n.invert;
        }
        case 3: {  // plane::in, cgs.cpp:92-98
            const double nc = r[0] * n.p[0] + r[1] * n.p[1] + r[2] * n.p[2];
            return nc > n.p[3] ? 0 : 1;
        }
        case 4: {  // cylinder::in, cgs.cpp:110-118
            for (int i = 0; i < 3; ++i) d[i] = r[i] - n.p[i];
            matvec(n.M, d, rel);
            if (rel[2] < 0 || rel[2] > n.p[3]) return n.invert;
            const double crs = rel[0] * rel[0] + rel[1] * rel[1];
            if (crs * n.p[3] > (n.p[6] - n.p[4]) * rel[2] + n.p[5]) return n.invert;
            return 1 - n.invert;
        }
    }
    return -1;
}

// composite_object::in (cgs.cpp:422-447) without recursion: explicit stack of frames.
__device__ int tree_in(int root, const double *r0) {
    struct Frame { int node; int stage; int l; double r[3]; };
    Frame st[SJ_MAX_DEPTH];
    int sp = 0;
    st[0].node = root; st[0].stage = 0; st[0].l = 0;
    matvec(c_nodes[root].M, r0, st[0].r);
    int ret = 0;
    while (sp >= 0) {
        Frame &f = st[sp];
        const sj_csg_node &n = c_nodes[f.node];
        if (f.stage == 0) {
            if (n.cmb == 3) { ret = 0; --sp; continue; }
            if (n.child0 < 0 && n.child1 < 0) { ret = n.invert; --sp; continue; }
            f.stage = 1;
            const int ch = n.child0 >= 0 ? n.child0 : n.child1;
            const sj_csg_node &cn = c_nodes[ch];
            if (cn.type == 0) {
                if (sp + 1 >= SJ_MAX_DEPTH) { ret = 0; --sp; continue; }
                st[sp + 1].node = ch; st[sp + 1].stage = 0; st[sp + 1].l = 0;
                matvec(cn.M, f.r, st[sp + 1].r);
                ++sp; continue;
            }
            ret = prim_in(cn, f.r);
            if (ret < 0) ret = n.invert;  // call_child_in fallthrough
        }
        if (f.stage == 1) {
            // ret holds the first evaluated child
            if (n.child0 < 0 || n.child1 < 0) { ret = n.invert ^ ret; --sp; continue; }
            f.l = ret; f.stage = 2;
            const sj_csg_node &cn = c_nodes[n.child1];
            if (cn.type == 0) {
                if (sp + 1 >= SJ_MAX_DEPTH) { ret = 0; --sp; continue; }
                st[sp + 1].node = n.child1; st[sp + 1].stage = 0; st[sp + 1].l = 0;
                matvec(cn.M, f.r, st[sp + 1].r);
                ++sp; continue;
            }
            ret = prim_in(cn, f.r);
            if (ret < 0) ret = n.invert;
        }
        if (f.stage == 2) {
            if (n.cmb == 0) ret = n.invert ^ (f.l | ret);
            else if (n.cmb == 1 || n.cmb == 2) ret = n.invert ^ (f.l & ret);
            else ret = n.invert;
            --sp; continue;
        }
    }
    return ret;
}

// coordinate of half-pixel index m; kind 0: h = m*(0.5*inva); kind 1: centre of the pixel volume
__device__ __forceinline__ double yee_coord(int m, double half, int kind) {
    const double h = m * half;
    if (kind == 0) return h;
    return ((h - half) + (h + half)) * 0.5;
}

__global__ void raster_kernel(uint8_t *out, int comp, int n_roots, int n0, int n1, int pitch, long long plane, int kz0,
                              int nzl, int n2, double half, int kind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kl = blockIdx.z;
    const int k = kz0 - 1 + kl;
    if (i > n0 || j > n1 || k < 0 || k > n2) return;
    double r[3];
    r[0] = yee_coord(2 * i + (comp == 0), half, kind);
    r[1] = yee_coord(2 * j + (comp == 1), half, kind);
    r[2] = yee_coord(2 * k + (comp == 2), half, kind);
    unsigned m = 0;
    for (int q = 0; q < n_roots; ++q)
        if (tree_in(c_roots[q], r)) m |= 1u << q;
    out[(long long)kl * plane + (long long)j * pitch + i] = (uint8_t)m;
}

// ---- stochastic boundary smoothing (smooth_n > 0) ---------------------------------------------------------------
// in_bound (disp.cpp:264-283) adds, for every region, the inside tests at 8 * smooth_n fixed offsets around the point:
// this_ret = in(r) + sum_j in(r + delta_j).  The kernel returns these integer sums (one byte per region, up to four
// regions packed in a 32-bit key) plus the plain inside bits; the host turns the distinct keys into the material table.
#define SJ_MAX_SMOOTH_PTS 248
__constant__ double c_smooth[3 * SJ_MAX_SMOOTH_PTS];

__global__ void raster_counts_kernel(uint32_t *keys, uint8_t *mask, int comp, int n_roots, int n_pts, int n0, int n1, int pitch,
                                     long long plane, int kz0, int nzl, int n2, double half) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kl = blockIdx.z;
    const int k = kz0 - 1 + kl;
    if (i > n0 || j > n1 || k < 0 || k > n2) return;
    const double r_x = yee_coord(2 * i + (comp == 0), half, 0);
    const double r_y = yee_coord(2 * j + (comp == 1), half, 0);
    const double r_z = yee_coord(2 * k + (comp == 2), half, 0);
    uint32_t key = 0;
    unsigned m = 0;
    for (int q = 0; q < n_roots; ++q) {
        double r[3] = {r_x, r_y, r_z};
        int cnt = tree_in(c_roots[q], r);
        if (cnt) m |= 1u << q;
        for (int s = 0; s < n_pts; ++s) {
            r[0] = r_x + c_smooth[3 * s]; r[1] = r_y + c_smooth[3 * s + 1]; r[2] = r_z + c_smooth[3 * s + 2];
            cnt += tree_in(c_roots[q], r);
        }
        key |= (uint32_t)cnt << (8 * q);
    }
    const long long x = (long long)kl * plane + (long long)j * pitch + i;
    keys[x] = key;
    mask[x] = (uint8_t)m;
}

// generate_smooth_pts (disp.cpp:56-112).  The reference draws the offsets with the C++ standard library's seed_seq /
// mt19937 / uniform_real_distribution / normal_distribution; so does this function, in the same call order, so the
// numbers are the reference's wherever both are linked against the same libstdc++ (oracle/csg_oracle.c restates the
// algorithms explicitly and the tests compare the two).  seed = DEF_SEED 0xd9a28bf3 through the reference's masks.
static std::vector<double> smooth_offsets(int smooth_n, double smooth_rad) {
    const uint64_t seed = 0xd9a28bf3ull;
    std::seed_seq seeder{(uint32_t)(seed & 0x0000ffff), (uint32_t)((seed & 0xffff0000) >> 32)};
    std::mt19937 gen(seeder);
    std::normal_distribution<double> gaussian(0, smooth_rad);
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    std::vector<double> pts((size_t)24 * smooth_n);
    for (int i = 0; i < smooth_n; ++i) {
        const double theta_inv = unif(gen);
        const double cos_theta = 1 - 2 * theta_inv;
        const double sin_theta = 2 * sqrt(theta_inv * (1 - theta_inv));
        const double phi = 2 * M_PI * unif(gen);
        const double r = gaussian(gen);
        const double x = r * sin_theta * cos(phi);
        const double y = r * sin_theta * cos(phi);      // sic (disp.cpp:92): cos, like x
        const double z = r * cos_theta;
        int j = 0;
        for (int xf = -1; xf < 2; xf += 2)
            for (int yf = -1; yf < 2; yf += 2)
                for (int zf = -1; zf < 2; zf += 2, ++j) {
                    pts[24 * i + 3 * j] = x * xf; pts[24 * i + 3 * j + 1] = y * yf; pts[24 * i + 3 * j + 2] = z * zf;
                }
    }
    return pts;
}

static int raster_smooth(sj_sim *s, double ambient_eps, int n_regions, const sj_region *regions, int smooth_n, double smooth_rad) {
    if (n_regions > 4) { s->err = "smooth_n > 0 supports at most 4 regions"; return SJ_ERR_UNSUPPORTED; }
    if (8 * smooth_n > SJ_MAX_SMOOTH_PTS) { s->err = "smooth_n > 31 is not supported"; return SJ_ERR_UNSUPPORTED; }
    const std::vector<double> pts = smooth_offsets(smooth_n, smooth_rad);
    const int n_pts = 8 * smooth_n;
    RCK(cudaMemcpyToSymbol(c_smooth, pts.data(), sizeof(double) * pts.size()));
    const double half = 0.5 * s->inva;
    const size_t npt = (size_t)s->set_stride;
    uint32_t *dkeys = NULL;
    RCK(cudaMalloc((void **)&dkeys, npt * sizeof(uint32_t)));
    std::vector<uint32_t> keys[3];
    for (int c = 0; c < 3; ++c) {
        if (!s->masks[c]) RCK(cudaMalloc((void **)&s->masks[c], npt));
        RCK(cudaMemsetAsync(s->masks[c], 0, npt, s->stream));
        RCK(cudaMemsetAsync(dkeys, 0, npt * sizeof(uint32_t), s->stream));
        dim3 blk(128), grd((s->g.n[0] + 1 + 127) / 128, s->g.n[1] + 1, s->nzl);
        raster_counts_kernel<<<grd, blk, 0, s->stream>>>(dkeys, s->masks[c], c, n_regions, n_pts, s->g.n[0], s->g.n[1], s->pitch,
                                                        s->plane, s->kz0, s->nzl, s->g.n[2], half);
        s->launches++;
        keys[c].resize(npt);
        RCK(cudaMemcpyAsync(keys[c].data(), dkeys, npt * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        RCK(cudaStreamSynchronize(s->stream));
    }
    cudaFree(dkeys);
    RCK(cudaGetLastError());
    // distinct count tuples -> material ids (ascending key order; key 0 = outside every region, also the padding)
    std::map<uint32_t, int> ids;
    ids[0] = 0;
    for (int c = 0; c < 3; ++c) for (size_t i = 0; i < npt; ++i) ids.emplace(keys[c][i], 0);
    if (ids.size() > 256) { s->err = "smoothing produced more than 256 distinct materials"; return SJ_ERR_UNSUPPORTED; }
    std::vector<sj_material> mats(ids.size());
    int next = 0;
    for (auto &kv : ids) {
        kv.second = next;
        sj_material &M = mats[next++];
        memset(&M, 0, sizeof M);
        double ret = 0.0;
        if (n_regions == 0) ret = ambient_eps;
        for (int r = 0; r < n_regions; ++r) {
            const double this_ret = (double)((kv.first >> (8 * r)) & 0xffu);
            ret += ambient_eps + (regions[r].eps - ambient_eps) * this_ret / (unsigned)(n_pts + 1);
            if (this_ret != 0)
                for (int q = 0; q < regions[r].n_poles; ++q) {
                    if (M.n_poles >= SJ_MAX_POLES) { s->err = "more than SJ_MAX_POLES poles overlap at one point"; return SJ_ERR_UNSUPPORTED; }
                    sj_pole pl = regions[r].poles[q];
                    pl.sigma = 0.0 + (pl.sigma - 0.0) * this_ret / (unsigned)(n_pts + 1);   // scale_func has def_ret = 0 (disp.cpp:545)
                    M.poles[M.n_poles++] = pl;
                }
        }
        M.eps_inf = ret;
    }
    s->mats = mats;
    std::vector<uint8_t> idb(npt);
    for (int c = 0; c < 3; ++c) {
        for (size_t i = 0; i < npt; ++i) idb[i] = (uint8_t)ids[keys[c][i]];
        RCK(cudaMemcpy(s->mat[c], idb.data(), npt, cudaMemcpyHostToDevice));
    }
    s->smoothed = true;
    return sj_finish_materials(s);
}

__global__ void or_masks(uint8_t *dst, const uint8_t *a, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] |= a[t];
}

// Sample coordinate: the Yee point itself, h = m * (0.5 * inva) (meep grid_volume::operator[] /
// IVEC_LOOP_LOC), for eps_inf and for the susceptibility sigma alike.  (meep's set_chi1inv
// evaluates eps at the centre of the pixel volume around the point, which can differ from h in
// the last bit; kind = 1 reproduces that form and is kept for experiments -- see DESIGN.md.)
int sj_raster_launch(sj_sim *s, double ambient_eps, int n_nodes, const sj_csg_node *nodes, int n_regions,
                     const sj_region *regions, int smooth_n, double smooth_rad) {
    if (n_nodes > SJ_MAX_NODES || n_regions > 8 || n_nodes < 0 || n_regions < 0) { s->err = "scene too large for the rasterizer"; return SJ_ERR_ARG; }
    int roots[8] = {0};
    for (int q = 0; q < n_regions; ++q) {
        if (regions[q].root < 0 || regions[q].root >= n_nodes || regions[q].n_poles < 0 || regions[q].n_poles > SJ_MAX_POLES) { s->err = "bad region"; return SJ_ERR_ARG; }
        roots[q] = regions[q].root;
    }
    if (n_nodes) RCK(cudaMemcpyToSymbol(c_nodes, nodes, sizeof(sj_csg_node) * n_nodes));
    RCK(cudaMemcpyToSymbol(c_roots, roots, sizeof roots));
    s->smoothed = false;
    if (smooth_n > 0) return raster_smooth(s, ambient_eps, n_regions, regions, smooth_n, smooth_rad);
    const double half = 0.5 * s->inva;
    for (int c = 0; c < 3; ++c) {
        if (!s->masks[c]) RCK(cudaMalloc((void **)&s->masks[c], (size_t)s->set_stride));
        RCK(cudaMemsetAsync(s->masks[c], 0, (size_t)s->set_stride, s->stream));
        dim3 blk(128), grd((s->g.n[0] + 1 + 127) / 128, s->g.n[1] + 1, s->nzl);
        raster_kernel<<<grd, blk, 0, s->stream>>>(s->masks[c], c, n_regions, s->g.n[0], s->g.n[1], s->pitch, s->plane,
                                                 s->kz0, s->nzl, s->g.n[2], half, 0);
        s->launches++;
    }
    RCK(cudaGetLastError());
    // Material table: material id == sigma-region mask (2^n_regions entries).  eps_inf follows
    // cgs_material_function::in_bound: sum over regions of def + (s_r - def) * in_r
    const int nm = 1 << n_regions;
    std::vector<sj_material> mats(nm);
    for (int m = 0; m < nm; ++m) {
        sj_material &M = mats[m];
        memset(&M, 0, sizeof M);
        double ret = 0.0;
        if (n_regions == 0) ret = ambient_eps;
        for (int r = 0; r < n_regions; ++r) {
            const double in = (m >> r) & 1;
            ret += ambient_eps + (regions[r].eps - ambient_eps) * in / (0 + 1);
            if (in != 0)
                for (int q = 0; q < regions[r].n_poles; ++q) {
                    if (M.n_poles >= SJ_MAX_POLES) { s->err = "more than SJ_MAX_POLES poles overlap at one point"; return SJ_ERR_UNSUPPORTED; }
                    M.poles[M.n_poles++] = regions[r].poles[q];
                }
        }
        M.eps_inf = ret;
    }
    s->mats = mats;
    for (int c = 0; c < 3; ++c) RCK(cudaMemcpyAsync(s->mat[c], s->masks[c], (size_t)s->set_stride, cudaMemcpyDeviceToDevice, s->stream));
    RCK(cudaStreamSynchronize(s->stream));
    return sj_finish_materials(s);
}

extern "C" int sj_rasterize(sj_sim *s, double ambient_eps, int32_t n_nodes, const sj_csg_node *nodes, int32_t n_regions,
                            const sj_region *regions) {
    if (!s || (n_nodes && !nodes) || (n_regions && !regions)) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    return sj_raster_launch(s, ambient_eps, n_nodes, nodes, n_regions, regions, 0, 0.0);
}

extern "C" int sj_rasterize_smooth(sj_sim *s, double ambient_eps, int32_t n_nodes, const sj_csg_node *nodes, int32_t n_regions,
                                   const sj_region *regions, int32_t smooth_n, double smooth_rad) {
    if (!s || (n_nodes && !nodes) || (n_regions && !regions) || smooth_n < 0) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    return sj_raster_launch(s, ambient_eps, n_nodes, nodes, n_regions, regions, smooth_n, smooth_rad);
}

// material id of every owned Yee point of E component comp, as an index into sj_get_material_table's order
extern "C" int sj_get_material_ids(sj_sim *s, int comp, uint8_t *out) {
    if (!s || comp < 0 || comp > 2 || !out) return SJ_ERR_ARG;
    cudaSetDevice(s->g.device);
    const int nx1 = s->g.n[0] + 1, ny1 = s->g.n[1] + 1;
    RCK(cudaStreamSynchronize(s->stream));
    for (int k = s->kz0; k < s->kz1; ++k)
        RCK(cudaMemcpy2D(out + (size_t)(k - s->kz0) * nx1 * ny1, nx1, s->mat[comp] + (size_t)(k - s->kz0 + 1) * s->plane, s->pitch,
                         nx1, ny1, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < (size_t)nx1 * ny1 * (s->kz1 - s->kz0); ++i) out[i] = s->lut_inv[out[i]];
    return SJ_OK;
}
