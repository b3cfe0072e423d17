// sj_internal.h -- shared declarations of the sim_juncs_b200 engine (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/sim_juncs_b200.h"

#define SJ_MAX_SRC 16
#define SJ_MAX_MAT 256
#define SJ_N_PML_BOX 18         // regions of the PML shell: 2 z boxes x 5 + 2 y boxes x 3 + 2 x boxes
#define SJ_N_AUX 10

// ---- kernel parameter blocks (passed by value) ---------------------------------------------
template <typename T>
struct SrcDev {
    int comp;
    int lo[3], hi[3];     // inclusive index ranges of the component grid (global indices)
    const T *w[3];        // per-direction weights, w[d][idx - lo[d]]
    T amp_re, amp_im;     // user amplitude * a^(zero-size dims)
};

template <typename T>
struct KParams {
    int n[3];             // global cells
    int pitch;            // row pitch in elements
    int rows;             // n[1] + 1
    long long plane;      // pitch * rows
    long long set_stride; // plane * nzl
    int kz0;              // global k of local plane 1 (local plane 0 is the lower halo)
    int nzl;              // local planes including both halos
    int n_sets;
    T *F;                 // all six components in one allocation: E[c] = F + c*fcs, H[c] = F + (3+c)*fcs
    long long fcs;        // component stride = set_stride * n_sets
    T *E[3];
    T *H[3];
    const uint8_t *mat[3];
    T *Pall;              // polarisation: [parity][slot][comp][set][planes p_k0 .. p_k0 + p_nzp) of the slab] in one allocation
    long long p_comp_stride;
    long long p_set_stride;   // p_nzp * plane: only the local planes [p_k0, p_k0 + p_nzp) that hold pole materials are stored
    int p_k0, p_nzp;
    int n_slots;
    int np_thr[SJ_MAX_POLES]; // material ids >= np_thr[s] have more than s poles (table sorted by pole count)
    const T *mt_chi;      // [SJ_MAX_MAT] 1/eps_inf
    const T *mt_eps;      // [SJ_MAX_MAT] eps_inf
    const int *mt_np;     // [SJ_MAX_MAT] number of poles
    int first_disp;       // material ids >= first_disp have poles (table is sorted that way)
    const T *mt_coef;     // [SJ_MAX_MAT][SJ_MAX_POLES][3] a1,a2,a3
    const T *sig[3];      // PML sigma*dt/2 per half-pixel index, length 2n+2
    const T *siginv[3];   // 1/(1+sig), meep's siginv table (kappa == 1)
    T courant;
    int n_src;
    int src_klo[SJ_MAX_SRC], src_khi[SJ_MAX_SRC];   // z index range of every source (block-uniform plane test)
    const SrcDev<T> *srcd;                          // [n_src] descriptors in device memory (read on source planes only)
    const T *drive;       // [step][src][set][2] = {S (integrated dipole), dt*current}
    const long long *step;// device step counter
};

// z-slab neighbours (sj_connect_*): where my boundary plane goes and the mailbox words that order the exchange.
// Protocol (step n = device step counter): after the E-pass of step n the upper slab has written its bottom E plane
// into my upper halo and set my flag_e = n + 1; my H-pass of step n waits for flag_e >= n before it touches that halo,
// writes its top H plane into the upper slab's lower halo and sets that slab's flag_h = n + 1, which the upper slab's
// E-pass of step n waits for.  The waits sit in the producer thread of the boundary items only.
template <typename T>
struct PeerLink {
    T *F;                             // neighbour's field allocation (peer-mapped); NULL = no neighbour connected
    long long fcs, set_stride;        // its component and set strides in elements
    int kl;                           // its local plane that receives my boundary plane (a halo plane of the neighbour)
    unsigned long long *flag;         // its mailbox word I set when the plane has landed
};
template <typename T>
struct SlabLinks {
    PeerLink<T> up, down;             // up: receives my top H plane; down: receives my bottom E plane
    unsigned long long *flag_e;       // mine: set by `up` (its bottom E plane is in my upper halo)
    unsigned long long *flag_h;       // mine: set by `down` (its top H plane is in my lower halo)
    unsigned int *done;               // [2] boundary-item completion counters (H-pass, E-pass)
    int n_bnd[2];                     // consumer-warp completions that finish the boundary plane of a pass
    int *err;                         // [1] set when a wait timed out
};
#define SJ_BND_FLAG 0x100             // WorkItem::shape bit: the item is the slab's boundary plane of this pass

// interior box + its fixed z-chunk grid (chunk c covers [k_lo + c*zchunk, k_lo + (c+1)*zchunk))
struct IntGeom {
    int i_lo, i_hi, j_lo, j_hi, k_lo, k_hi;
    int zchunk, nzc, c0;        // chunks launched: c0 .. c0+nzc-1
    const unsigned *flags;      // [chunk][tile_y][tile_x] material flags (e_interior)
};

template <typename T>
struct PmlBox {
    int lo[3], hi[3];     // global index box [lo,hi) ; k clipped to the local slab
    int bx, by;           // box extents in x, y
    int bpitch;           // row pitch of aux arrays
    long long bplane;     // bpitch * by
    long long bset;       // elements per set
    long long bcs;        // component stride of the box arrays (= bset * n_sets); X[c] = X[0] + c*bcs
    T *D[3], *B[3], *UD[3], *UB[3];
};

template <typename T>
struct PmlBoxSet { PmlBox<T> b[SJ_N_PML_BOX]; };

// one thread block of a PML tile kernel: a (tile_w x tile_h) column of box `box`, planes [kb,ke)
// box: PML box (-1 interior); [i0,i_hi) x [j0,j_hi): the tile clipped to its rectangle; mat: uniform material id
// (fast-path lists); kind: 0 general PML cell, 1/2/3 face tile with normal x/y/z
// shape: tile shape index of the TMA-staged kernels (sj_tma.cuh)
struct WorkItem { int box, set, i0, j0, kb, ke, mat, kind, i_hi, j_hi, shape, pad; };
struct ItemList { WorkItem *dev; int n; };

// ---- TMA-staged column kernels (sj_tma.cuh): tile shapes, tensor maps, per-block item schedules ----------------
#define SJ_TMA_MAX_SHAPES 12
#ifndef SJ_TMA_NT
#define SJ_TMA_NT 224              // consumer threads per block of the TMA-staged kernels (7 warps + 1 producer warp at 255 registers)
#endif
// nvx vectors per tile row, th tile rows, tw tile width and hp halo-box row pitch in elements; a thread owns vector
// t % nvx of rows t / nvx + s * ths, s < sub (sub = 2: tall tiles, two rows per thread -- a TMA box of twice the bytes costs
// the TMA unit the same ~170 cycles); hs / os: bytes of a halo box / an own-cell tile in the staging ring (128-byte multiples)
struct TShape { int nvx, th, tw, hp, sub, ths, hs, os; };
struct TmaList { WorkItem *items; int *first; int n_items, grid; double bytes; };   // first: the queue counters {next item, drained producers}
struct TmaState {
    int mode = 0;                 // bit 0: H-pass through the TMA kernels, bit 1: E-pass
    int nt = SJ_TMA_NT;           // consumer threads per block the shapes were sized for
    int n_shapes = 0;
    TShape shapes[SJ_TMA_MAX_SHAPES];
    void *maps = nullptr;         // CUtensorMap[n_shapes][SJ_TMAP_PER_SHAPE] in device memory
    std::vector<WorkItem> geo[2]; // geometry items of one field set: [H-pass | E-pass (before material classification)]
    int n_bnd[2] = {0, 0};        // boundary-plane items per field set in the two lists
    TmaList h[2] = {};            // H-pass: [class A (interior + faces) | general (edges, corners)]
    TmaList e[2][4] = {};         // E-pass: same split x material class (0: one non-dispersive material, 1: mixed,
                                  //         2 / 3: one material with 1 / 2 poles)
    // fused step (one launch, z wavefront; single-slab runs whose plane fits the L2 window): both passes' items in one queue
    TmaList f = {};
    int wave = 0;                 // planes per z chunk (0: no fused list)
    int n_chunks = 0;
    int *grp = nullptr;           // device: [need (n_chunks) | done (n_chunks) | epoch]
};

struct MonDev {
    int n_mon;
    int comp;
    const long long *idx; // [n_mon][8] local linear indices (-1 = skip)
    const double *w;      // [n_mon][8]
    double *series;       // [cap][n_mon][n_sets]
    int *flags;           // [0] = diverged
    const unsigned long long *wait_flag;   // z-slab runs: mailbox word to wait on before sampling (NULL otherwise)
};

// ---- host-side state -------------------------------------------------------------------
struct HostSource {
    int comp, integrated;
    int kind;                     // 0: gaussian_src_time_phase, 1: meep::continuous_src_time, 2: caller's waveform
    void (*fn)(void *, double, double *) = nullptr;   // kind 2: dipole(ctx, time, {re, im}), called on the host
    void *fn_ctx = nullptr;
    double last_time = 0;         // kind 2
    double t_start, t_end, slowness;   // kind 1
    double omega, width, phi, peak, cutoff;
    double amp_t_re, amp_t_im;    // 1/(-i omega)
    double amp_re, amp_im;
    int lo[3], hi[3];
    std::vector<double> w[3];
    std::vector<double> set_phase;
};

struct sj_sim {
    sj_grid g;
    int prec;                 // SJ_F64 / SJ_F32
    size_t esz;
    double inva, dt;
    int pitch, rows, nzl;
    long long plane, set_stride;
    int kz0, kz1;
    int lo[3], hi[3];         // interior box [lo,hi) in global indices (x aligned)
    std::vector<double> sig[3];

    void *F;                  // one allocation holding E[0..2], H[0..2]
    void *E[3], *H[3];        // device, [set][local plane][row][pitch]
    uint8_t *mat[3];
    uint8_t *masks[3];        // region masks from the rasterizer (same layout)
    bool smoothed = false;    // materials come from the smooth_n > 0 rasterizer (ids are not region masks)
    void *Pall;
    int p_alloc_k0 = -1, p_alloc_nzp = -1; double p_bytes = 0;
    int p_k0 = 1, p_nzp = 1;  // local planes [p_k0, p_k0 + p_nzp) have polarisation storage (the planes that hold pole materials)
    int n_slots;
    int np_thr[SJ_MAX_POLES];
    void *mt_chi, *mt_eps, *mt_coef; int *mt_np;
    void *sigd[3], *siginvd[3];
    unsigned *flags_int;
    std::vector<WorkItem> h_items[2][2];   // PML tiles: [general|face][wide|narrow]
    ItemList il_h[2][2];                   // H-pass PML lists, same indexing
    // E-pass lists by material class of the tile: 0 = one non-dispersive material, 1 = mixed (general path),
    // 2 / 3 = one material with exactly 1 / 2 poles (uniform-dispersive fast path).  interior; PML [general|face][wide|narrow]
    ItemList il_int[4], il_pml[2][2][4];
    int pml_lx, pml_lx_n;                  // lanes along x of the wide / narrow PML tiles
    int pml_v;                             // elements per thread in the PML kernels (full or half vector)
    int int_lx, int_zchunk;   // interior tiling: lanes along x per warp, planes per chunk
    int first_disp;
    bool present[256];        // material id (sorted order) occurs in the slab
    uint8_t lut_inv[256];     // material id -> id as given by the caller
    bool materials_set;
    void *src_dev;            // SrcDev<T>[SJ_MAX_SRC] on the device
    std::vector<sj_material> mats;         // as given by the caller / rasterizer (id = caller id)
    std::vector<sj_material> mats_sorted;  // device order: non-dispersive first

    struct Box { int lo[3], hi[3]; int bx, by, bz, bpitch; int kind, narr; long long bplane, bset; void *base; void *D[3], *B[3], *UD[3], *UB[3]; };
    std::vector<Box> boxes;

    std::vector<HostSource> srcs;
    void *srcw[SJ_MAX_SRC][3];
    void *drive; long long drive_steps; bool drive_dirty;

    int n_mon, mon_comp;
    std::vector<double> mon_xyz;
    long long *mon_idx; double *mon_w; double *series; int series_cap; int n_samples;
    std::vector<char> mon_owned;
    int *flags;

    long long *step_dev;
    long long steps_done;
    cudaStream_t stream;
    cudaEvent_t ev_a, ev_b;
    cudaStream_t aux[SJ_N_AUX];           // side streams: the independent kernels of a half-pass run concurrently
    cudaEvent_t ev_fork, ev_join[SJ_N_AUX];
    int fan_next, fan_int; cudaStream_t fan_main; bool fan_on; int n_aux;
    // SJ_TRACE=1: per-launch CUDA events of one traced step (debug timeline, printed by sj_trace_dump)
    bool trace_on; std::vector<cudaEvent_t> tr_ev; std::vector<std::string> tr_name; cudaEvent_t tr_origin;
    // one full time step (H-pass, E-pass, tick; side-stream fan-out included) captured as a CUDA graph and
    // replayed by sj_run; dropped whenever a pointer or a work list the kernels receive changes
    cudaGraphExec_t step_graph; long long graph_launches;
    long long launches;
    double pole_points;       // sum over E component points of n_poles (owned slab)
    double pole_points_int;   // same, restricted to the interior-kernel box
    double pml_cells;
    double pml_bytes = 0;     // bytes of UPML auxiliary storage
    double h2d_bytes = 0;     // source drive table bytes uploaded so far
    TmaState tma;
    // z-slab neighbours: raw peer pointers (same process) or CUDA-IPC mappings (other processes)
    struct Peer { void *F = nullptr; void *sync = nullptr; long long fcs = 0, set_stride = 0; int nzl = 0; bool ipc = false; };
    Peer peer_up, peer_down;
    unsigned long long *sync_dev = nullptr;   // [0] flag_e, [1] flag_h, [2] done counters (2 x u32), [3] spare
    bool trace_reg = false;   // force the register kernels (profiling the old path)
    std::string err;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            s->err = b_;                                                                           \
            return SJ_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

int sj_finish_materials(sj_sim *s);
void sj_interior_geom(const sj_sim *s, int k_begin, int k_end, IntGeom &g, dim3 &grd);
// sj_launch_f64.cu / sj_launch_f32.cu: one half-pass over the planes [k0, k1) on stream st; kernel-family timings
int sj_launch_pass_f64(sj_sim *s, int which, int k0, int k1, cudaStream_t st);
int sj_launch_pass_f32(sj_sim *s, int which, int k0, int k1, cudaStream_t st);
int sj_profile_f64(sj_sim *s, int reps, double out[4]);
int sj_profile_f32(sj_sim *s, int reps, double out[4]);
// sj_tma_host.cu / sj_tma_f64.cu / sj_tma_f32.cu
int sj_classify_items(sj_sim *s, const std::vector<WorkItem> &items, int tile_w, int tile_h, std::vector<WorkItem> (&out)[4]);
int sj_tma_build_geometry(sj_sim *s);          // shapes + geometry items + H-pass schedules (once, at create)
int sj_tma_build_materials(sj_sim *s);         // tensor maps + E-pass schedules (after every material upload)
void sj_tma_free(sj_sim *s);
int sj_tma_pass_f64(sj_sim *s, int which, cudaStream_t st, bool tick);
int sj_tma_pass_f32(sj_sim *s, int which, cudaStream_t st, bool tick);
// sj_raster.cu
int sj_raster_launch(sj_sim *s, double ambient_eps, int n_nodes, const sj_csg_node *nodes, int n_regions,
                     const sj_region *regions, int smooth_n, double smooth_rad);
