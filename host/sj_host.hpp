// sj_host.hpp -- C++ host for libsimjuncs_b200: the reference's `bound_geom` (src/disp.hpp:142-200)
// re-hosted on the C ABI.  It keeps the reference's front end untouched: parse_settings / params.conf
// (src/argparse.h, header-only, included as is) and the CGS scene language (src/cgs*.{hpp,cpp},
// compiled in place by host/Makefile).  Only the meep members are replaced by one sj_sim handle.
#pragma once
#include <complex>
#include <string>
#include <vector>

#include "argparse.h"             // reference header (parse_settings, correct_defaults, ...)
#include "cgs.hpp"                // reference header (scene, context, value, composite_object)
#include "sim_juncs_b200.h"

#define SJ_LIGHT_SPEED 0.299792458
#define SJ_THICK_SCALE 1.0

struct sj_pole_raw { double omega_0, gamma, sigma; bool use_denom; };

enum sj_src_type { SJ_SRC_GAUSSIAN, SJ_SRC_CONTINUOUS };
struct sj_source_info {             // source_info, src/disp.hpp:98-110 (raw .geom units: um, fs)
    sj_src_type type;
    int component;                  // 0..5 = Ex Ey Ez Hx Hy Hz
    double wavelen, width, phase, start_time, end_time, amplitude;
    bool ok;
    sj_source_info(value info);
};

struct sj_vec3 { double x, y, z; };

// flatten a composite_object tree into C-ABI nodes (host/sj_flatten.cpp)
int sj_flatten_tree(composite_object *root, std::vector<sj_csg_node> &nodes);

class sj_bound_geom {
public:
    // n_gpus > 1: the simulation is cut into z-slabs of equal bytes per step, one per GPU of this box (the reference's
    // counterpart is meep under MPI, src/main.cpp:20); results equal the single-GPU run bit for bit
    sj_bound_geom(const parse_settings &s, parse_ercode *ercode = NULL, int precision = SJ_F64, int n_sets = 2,
                  int integrated = 1, int n_gpus = 1);
    ~sj_bound_geom();
    std::vector<sj_vec3> get_monitor_locs() { return monitor_locs; }
    std::vector<std::vector<std::complex<double> > > get_field_times() { return field_times; }
    std::vector<sj_source_info> get_sources() { return sources; }
    size_t get_n_monitor_clusters() const { return monitor_clusters.size(); }
    int run(const char *fname_prefix);
    int save_field_times(const char *fname_prefix);       // field_samples.h5 (own encoder, sj_hdf5.hpp) + field_samples.npz
    double fs_to_meep_time(double t) const { return t * SJ_LIGHT_SPEED * um_scale; }
    double meep_time_to_fs(double t) const { return t / (SJ_LIGHT_SPEED * um_scale); }
    std::vector<sj_pole_raw> parse_susceptibilities(value val, int *er);
    double raster_ms, run_s;
    unsigned n_t_pts;
    double ttot;

    scene problem;

private:
    sj_sim *sim;                      // the bottom slab (the only one on a single GPU)
    std::vector<sj_sim *> sims;       // all z-slabs, bottom to top
    std::vector<sj_source_info> sources;
    std::vector<sj_vec3> monitor_locs;
    std::vector<size_t> monitor_clusters;
    std::vector<std::vector<std::complex<double> > > field_times;
    double um_scale, post_source_t;
    unsigned save_span;
    int n_sets;
    bool dump_raw;
    int n_cells[3];
};

context sj_context_from_settings(const parse_settings &args);
