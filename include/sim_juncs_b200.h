/*
 * sim_juncs_b200.h -- C ABI of libsimjuncs_b200.so, the B200-native FDTD engine that replaces
 * libmeep underneath sim_juncs' `bound_geom` (reference src/disp.hpp:142-200).
 *
 * The reference has no plugin API: its seam into the engine is the set of meep calls made by
 * src/disp.cpp.  Each entry point below names the reference call site it stands in for.  All
 * pointers are plain host pointers unless the name says `dev`; the library copies what it needs
 * during the call and owns all device memory behind the opaque handle.  Every function returns
 * 0 on success or a negative sj_status; sj_last_error() gives the text.  Nothing throws or exits.
 *
 * Conventions (identical to the reference's use of meep): lengths in meep units, c = 1,
 * dx = 1/a, dt = courant/a, Yee grid anchored at the low corner, metallic outer wall behind the
 * PML.  A field component array has (n[0]+1) x (n[1]+1) x (n[2]+1) points, x fastest.
 */
#ifndef SIM_JUNCS_B200_H
#define SIM_JUNCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sj_sim sj_sim;

typedef enum {
    SJ_OK = 0,
    SJ_ERR_ARG = -1,       /* bad argument */
    SJ_ERR_CUDA = -2,      /* CUDA runtime failure (message in sj_last_error) */
    SJ_ERR_STATE = -3,     /* call order violated (e.g. stepping before materials are set) */
    SJ_ERR_NOMEM = -4,
    SJ_ERR_DIVERGED = -5,  /* a monitor sample exceeded 1000 or went non-finite (disp.cpp:727) */
    SJ_ERR_UNSUPPORTED = -6
} sj_status;

enum { SJ_EX = 0, SJ_EY = 1, SJ_EZ = 2, SJ_HX = 3, SJ_HY = 4, SJ_HZ = 5 };
enum { SJ_F64 = 0, SJ_F32 = 1 };

/* Grid + engine configuration.  Replaces meep::vol3d(L,L,L,a) + meep::structure(vol, eps,
 * meep::pml(thickness)) + meep::fields(structure)   (reference src/disp.cpp:505-509,527,560). */
typedef struct {
    int32_t n[3];          /* cells per direction: (int)(L*a + 0.5), meep vol3d                    */
    double a;              /* resolution (pixels per length unit), argparse.h:171-191 rounding     */
    double courant;        /* 0.5: params.conf `courant` is parsed but never forwarded (fact 0.9)  */
    double pml_thickness;  /* meep::pml(thickness) on all six faces; 0 disables                    */
    double pml_R;          /* asymptotic reflection, meep default 1e-15                            */
    int32_t precision;     /* SJ_F64 | SJ_F32 storage + arithmetic of fields                       */
    int32_t n_sets;        /* independent field sets sharing the materials: 2 = meep complex
                              fields (Re, Im), 1 = real fields, >2 = phase batch                   */
    int32_t kz0, kz1;      /* owned z-planes [kz0,kz1) of the n[2]+1 planes (z-slab); 0,0 = all    */
    int32_t device;        /* CUDA device ordinal, -1 = current                                    */
} sj_grid;

/* One Drude-Lorentz pole = one meep::lorentzian_susceptibility (src/disp.cpp:536-546), already in
 * meep units: omega0 = file/um_scale, gamma = file/um_scale, sigma = file/thickness. */
typedef struct {
    double omega0, gamma, sigma;
    int32_t drude;         /* !use_denom -> no_omega_0_denominator */
    int32_t pad;
} sj_pole;

#define SJ_MAX_POLES 4
/* A material = one distinct combination of scene regions at a Yee point. */
typedef struct {
    double eps_inf;        /* cgs_material_function::in_bound sum (src/disp.cpp:264-283) */
    int32_t n_poles;
    int32_t pad;
    sj_pole poles[SJ_MAX_POLES];
} sj_material;

/* Flattened CSG node (reference classes object/sphere/box/plane/cylinder/composite_object,
 * src/cgs.hpp:36-157); layout shared with oracle/csg_oracle.c. */
typedef struct {
    int32_t type;          /* 0 composite 1 sphere 2 box 3 plane 4 cylinder 5 undefined */
    int32_t invert;
    int32_t cmb;           /* combine_type: 0 union 1 intersect 2 difference 3 noop */
    int32_t child0, child1;/* node indices, -1 = NULL */
    int32_t pad;
    double M[9];           /* trans_mat */
    double p[7];           /* primitive parameters, see csg_oracle.c */
} sj_csg_node;

/* One scene region (a `Composite(...)` root) with the metadata the engine consumes. */
typedef struct {
    int32_t root;          /* index of the root node */
    int32_t n_poles;
    double eps;            /* metadata `eps`, or ambient when absent (add_region, disp.cpp:249-262) */
    sj_pole poles[SJ_MAX_POLES];
} sj_region;

/* ---- lifetime ------------------------------------------------------------------------- */
int sj_create(const sj_grid *grid, sj_sim **out);
void sj_destroy(sj_sim *sim);
const char *sj_last_error(const sj_sim *sim);   /* sim may be NULL: last create() failure */
int sj_version(void);

/* ---- materials -------------------------------------------------------------------------
 * Either hand over per-component material indices (host arrays over the GLOBAL grid, x fastest,
 * one byte per Yee point of Ex / Ey / Ez) ...                                                */
int sj_set_materials(sj_sim *sim, int32_t n_mat, const sj_material *mats, const uint8_t *idx_ex,
                     const uint8_t *idx_ey, const uint8_t *idx_ez);
/* ... or let the GPU rasterize the scene (replaces the 3N^3 + 3N^3/pole virtual in_bound() calls
 * meep makes at setup: chi1p1 -> src/disp.cpp:285-288, sigma_row -> :303-306).  Regions are
 * summed exactly like in_bound(); the material table is built from the distinct region masks.  */
int sj_rasterize(sj_sim *sim, double ambient_eps, int32_t n_nodes, const sj_csg_node *nodes,
                 int32_t n_regions, const sj_region *regions);
/* The same with the reference's stochastic boundary smoothing (params.conf smooth_n / smooth_rad; cgs_material_function::
 * generate_smooth_pts + in_bound, src/disp.cpp:56-112, 264-283): every inside test becomes the sum over the point and
 * 8 * smooth_n fixed offsets drawn from mt19937(seed_seq{0x8bf3, 0}); eps_inf and every pole's sigma take one of
 * 8 * smooth_n + 2 levels per region.  A material is then one distinct tuple of per-region sums (at most 256 tuples,
 * 4 regions, smooth_n <= 31; SJ_ERR_UNSUPPORTED beyond).  smooth_n = 0 is sj_rasterize.                              */
int sj_rasterize_smooth(sj_sim *sim, double ambient_eps, int32_t n_nodes, const sj_csg_node *nodes,
                        int32_t n_regions, const sj_region *regions, int32_t smooth_n, double smooth_rad);
/* Material id (index into sj_get_material_table) of every owned Yee point of E component comp. */
int sj_get_material_ids(sj_sim *sim, int comp, uint8_t *out);
/* Read back what sj_rasterize produced (tests, eps-*.h5 dump):  region bit masks per E component
 * over the owned slab planes [kz0,kz1) and the material table. */
int sj_get_region_masks(sj_sim *sim, int comp, uint8_t *out);
int sj_get_material_table(sj_sim *sim, int32_t *n_mat, sj_material *out, int32_t cap);

/* ---- sources ---------------------------------------------------------------------------
 * fields.add_volume_source(component, gaussian_src_time_phase(f, w, phase, t0, t1), volume, amp)
 * (src/disp.cpp:603-619; waveform src/disp.cpp:378-400).  Times in meep units.  `integrated`
 * selects meep's src_time::is_integrated (the reference inherits the base-class default, 1).
 * set_phase (may be NULL) gives an extra phase per field set for phase batches (n_sets > 2);
 * with n_sets <= 2 set 0 takes the real and set 1 the imaginary part of the complex drive.   */
int sj_add_gaussian_source(sj_sim *sim, int comp, const double lo[3], const double hi[3],
                           double amp_re, double amp_im, double freq, double width, double phase,
                           double t_start, double t_end, int integrated, const double *set_phase);
/* CW_source -> meep::continuous_src_time(freq, width, t_start, t_end) (src/disp.cpp:615-619; meep's default
 * slowness 3.0): exp(-i w t)/(-i w) between t_start and t_end with tanh turn-on / turn-off of time
 * constant `width` (0 = abrupt).  t_end must be finite.                                                  */
int sj_add_cw_source(sj_sim *sim, int comp, const double lo[3], const double hi[3],
                     double amp_re, double amp_im, double freq, double width,
                     double t_start, double t_end, double slowness, int integrated, const double *set_phase);
/* fields.add_volume_source(c, src, volume, amp) for ANY meep::src_time subclass: `fn` is its dipole(time) (called on
 * the host, from the calling thread, while the drive table of the next steps is built; it must stay valid until
 * sj_destroy), last_time its last_time(), integrated its is_integrated.                                          */
typedef void (*sj_dipole_fn)(void *ctx, double time, double out_re_im[2]);
int sj_add_custom_source(sj_sim *sim, int comp, const double lo[3], const double hi[3],
                         double amp_re, double amp_im, sj_dipole_fn fn, void *ctx, double last_time,
                         int integrated, const double *set_phase);
double sj_last_source_time(const sj_sim *sim);     /* fields.last_source_time(), disp.cpp:625 */

/* ---- monitors: fields.get_field(component, loc) at fixed points (src/disp.cpp:724) ------- */
int sj_add_monitors(sj_sim *sim, int comp, int32_t n, const double *xyz);

/* fields.get_field(component, loc) at arbitrary points, at the current step (one small kernel + one copy; the monitor
 * list above is the fast path).  out: [n][n_sets] doubles, same interpolation and ownership rule as the monitors. */
int sj_sample_at(sj_sim *sim, int comp, int32_t n, const double *xyz, double *out);

/* ---- stepping --------------------------------------------------------------------------
 * sj_run = the body of bound_geom::run (src/disp.cpp:719-741): for i in [0,n_steps): if
 * (i % save_span == 0) sample every monitor; fields.step().  Samples stay on the device until
 * sj_read_monitors.  May be called repeatedly; the step counter and sample cursor continue.   */
int sj_run(sj_sim *sim, int64_t n_steps, int32_t save_span);
int sj_sync(sj_sim *sim);
int64_t sj_steps_done(const sj_sim *sim);
int32_t sj_n_samples(const sj_sim *sim);
double sj_dt(const sj_sim *sim);
/* out: [n_samples][n_monitors][n_sets] doubles (set 0 = Re, set 1 = Im in complex mode).
 * In a z-slab run each rank fills the monitors it owns and leaves the others 0. */
int sj_read_monitors(sj_sim *sim, double *out);
/* Spectra of the monitor series, computed on the device: the reference's fft() (src/data_utils.cpp:370-427) as
 * save_field_times calls it (src/disp.cpp:806) -- only the first T = 2^floor(log2 N) of the N samples enter, the phase
 * step is 2 pi / N of the full length, output in FFT order.  The complex series is (set_re, set_im); set_im = -1 for a
 * real series.  *n_freq = T; out (may be NULL to query T): [n_monitors][T]{re, im} doubles. */
int sj_read_spectra(sj_sim *sim, int32_t set_re, int32_t set_im, int32_t *n_freq, double *out);

/* Pulse parameters of every monitor series, on the device: the front half of the reference's post-processing class
 * `signal` (scripts/phases.py:94-157, 205-244, 289-357), which its phase sweeps evaluate in a Python loop over 1 600
 * series x phases -- arrival-sample guess, 2 rfft of the rolled series, magnitude-weighted centre frequency, the band
 * above CUT_ALPHA = 0.1 of the centre, unwrapped phase regressed on the frequency.  dt_sample: time between two samples
 * (any unit; frequencies come out in its inverse).  out: [n_sets][n_monitors][10] doubles = f0, f0 bin, arrival-sample
 * guess, t0_corr = -slope / 2 pi, phi_corr = fix_angle(intercept), slope, intercept, r value, f_min + 65536 f_max,
 * status (bit 0: the reference would low-pass this series first -- run the host pipeline on it; bit 1: centre frequency
 * above half the band; bit 2: band too narrow for a fit). */
int sj_extract_cep(sj_sim *sim, double dt_sample, double *out);

/* ---- z-slab halo exchange (one process per GPU; the caller moves the bytes) -------------
 * Fine-grained stepping used by the multi-GPU driver: half-passes restricted to local plane
 * ranges so boundary planes can be computed first and exchanged while the rest runs.         */
int sj_pass(sj_sim *sim, int which /*0 = H-pass, 1 = E-pass*/, int32_t k_begin, int32_t k_end,
            void *cuda_stream);
/* cuda_stream (here and below): the cudaStream_t the launches are ordered on.  NULL selects the simulation's own
 * non-blocking stream, which is NOT ordered with the legacy default stream (handle 0): a caller that moves the
 * halos with NCCL / MPI on its own stream must pass that stream. */
int sj_tick(sj_sim *sim, void *cuda_stream);        /* end of step: advance device step counter */
int sj_sample(sj_sim *sim, void *cuda_stream);      /* sample monitors now */
/* Device pointer + byte count of plane k (global index, may be a halo plane kz0-1 / kz1) of a
 * component array of field set `set`. */
int sj_plane_ptr(sj_sim *sim, int comp, int set, int32_t k, void **dev_ptr, size_t *bytes);

/* Single-process helper for two slabs that are stacked in z (lower.kz1 == upper.kz0), on the same or
 * on peer devices: which = 0 copies the top owned Hx,Hy planes of `lower` into the lower halo of
 * `upper` (after an H-pass); which = 1 copies the bottom owned Ex,Ey planes of `upper` into the upper
 * halo of `lower` (after an E-pass).  Stream-ordered on both simulations' streams. */
int sj_halo_exchange(sj_sim *lower, sj_sim *upper, int which);

/* ---- z-slabs inside the library (the reference's counterpart is meep's MPI chunking under `meep::initialize mpi`,
 * src/main.cpp:20) -------------------------------------------------------------------------------------------------
 * Every slab is one sj_sim created with its [kz0, kz1).  Once the neighbours are connected, sj_run / sj_run_group need
 * nothing else: the kernels that update a slab's boundary plane write it straight into the neighbour's halo over
 * NVLink (peer-mapped memory) and the slabs order themselves through mailbox words on the device; the step still
 * replays as one CUDA graph per slab.  Fields and monitors equal the single-GPU run bit for bit. */
typedef struct { unsigned char bytes[256]; } sj_peer_handle;
/* one process per GPU: export this slab's CUDA-IPC handle, hand it to the neighbours' processes (any byte transport),
 * connect with the handles of the slab below / above (NULL where there is none) -- before the first step. */
int sj_export_peer(sj_sim *sim, sj_peer_handle *out);
int sj_connect_peers(sj_sim *sim, const sj_peer_handle *lower, const sj_peer_handle *upper);
/* one process, several devices (or several slabs on one device): connect two stacked slabs directly */
int sj_connect_local(sj_sim *lower, sj_sim *upper);
/* sj_run for all slabs of a process (bottom to top), enqueued so that no slab's launch queue starves its neighbours */
int sj_run_group(sj_sim **sims, int32_t n_slabs, int64_t n_steps, int32_t save_span);

/* ---- inspection ------------------------------------------------------------------------- */
/* Copy a whole component (owned planes [kz0,kz1), dense (n0+1)(n1+1) rows) to the host as doubles. */
int sj_get_field(sj_sim *sim, int comp, int set, double *out);
/* sj_run bracketed by CUDA events on the simulation's stream; *ms = device time of the n_steps. */
int sj_run_timed(sj_sim *sim, int64_t n_steps, int32_t save_span, double *ms);
/* Average device time (ms, CUDA events, `reps` back-to-back launches each) of the step's kernels over the owned slab:
 * out[0] the H-pass kernel, out[1] the E-pass kernel (each one persistent TMA kernel covering interior and PML cells;
 * out[2] = out[3] = 0).  With SJ_TMA=0 (register kernels): H interior, E interior, H PML lists, E PML lists.
 * Advances nothing logically (call on a scratch simulation). */
int sj_profile_kernels(sj_sim *sim, int32_t reps, double out[4]);
/* out[0] owned cells, out[1] interior-kernel cells, out[2] PML-kernel cells, out[3] pole-points
 * (sum over E-component points of n_poles) in the slab, out[4] pole-points inside the interior
 * box, out[5] true PML cells (any sigma != 0). */
int sj_get_counts(sj_sim *sim, double out[6]);
/* Algorithmic bytes per step of every owned plane (the terms of sj_bytes_per_step plane by plane): the weights a z-slab
 * decomposition balances.  out: [kz1 - kz0] doubles. */
int sj_plane_costs(sj_sim *sim, double *out);
/* Device memory of this slab in bytes: out[0] E and H, out[1] polarisation (stored only over the planes that hold pole
 * materials), out[2] UPML auxiliaries (faces keep the normal D and B only), out[3] material indices, out[4] their sum,
 * out[5] number of planes with polarisation storage. */
int sj_memory(const sj_sim *sim, double out[6]);
/* Per-run statistics for bench.py: kernels launched so far and bytes of source drive table uploaded (host -> device). */
int sj_get_stats(const sj_sim *sim, int64_t *kernel_launches, double *h2d_bytes);
/* Algorithmic bytes one full step moves for this configuration. */
double sj_bytes_per_step(const sj_sim *sim);

#ifdef __cplusplus
}
#endif
#endif /* SIM_JUNCS_B200_H */
