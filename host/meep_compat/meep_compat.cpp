// host/meep_compat/meep_compat.cpp -- the meep API slice of meep.hpp on libsimjuncs_b200's C ABI.  See meep.hpp.
// Environment: SJ_PRECISION=f32 selects fp32 fields (default fp64, like meep); SJ_DEVICE=<ordinal>.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>

#include "meep.hpp"
#include "sim_juncs_b200.h"
#include "../sj_dumps.hpp"

namespace meep {

int verbosity = 1;

static void ck(int rc, sj_sim *sim, const char *what) {
    if (rc == SJ_OK || rc == SJ_ERR_DIVERGED) return;       // divergence is reported by the reference's own Re > 1000 check
    throw std::runtime_error(std::string(what) + ": " + sj_last_error(sim));
}

static grid_volume make_gv(ndim d, double a, int nx, int ny, int nz) {
    grid_volume g;
    g.dim = d; g.a = a; g.n[0] = nx; g.n[1] = ny; g.n[2] = nz;
    return g;
}
// meep vec.cpp: the number of pixels is (int)(size * a + 0.5)
grid_volume vol1d(double zsize, double a) { return make_gv(D1, a, 0, 0, (int)(zsize * a + 0.5)); }
grid_volume vol2d(double xsize, double ysize, double a) { return make_gv(D2, a, (int)(xsize * a + 0.5), (int)(ysize * a + 0.5), 0); }
grid_volume vol3d(double xsize, double ysize, double zsize, double a) {
    return make_gv(D3, a, (int)(xsize * a + 0.5), (int)(ysize * a + 0.5), (int)(zsize * a + 0.5));
}
boundary_region pml(double thickness) { return boundary_region(thickness); }

// meep::continuous_src_time::dipole as the engine evaluates it for sj_add_cw_source
std::complex<double> continuous_src_time::dipole(double time) const {
    float rtime = float(time);
    if (rtime < start_time || rtime > end_time) return 0.0;
    const double omega = 2 * M_PI * real(freq);
    const std::complex<double> amp = 1.0 / std::complex<double>(0, -omega);
    const std::complex<double> osc = std::polar(1.0, -omega * time) * amp;
    if (width == 0.0) return osc;
    const double ts = (time - start_time) / width - slowness;
    const double te = (end_time - time) / width - slowness;
    return osc * (1.0 + tanh(ts)) * (1.0 + tanh(te)) * 0.25;
}

// Yee point of E component c at array index (i, j, k): half-pixel index m = 2 i + (d == c), coordinate m * (0.5 / a),
// the sample coordinate of the engine's own rasterizer (sim_juncs_b200/csrc/sj_raster.cu)
template <typename F>
static void for_each_yee(const grid_volume &gv, int c, std::vector<double> &out, F f) {
    const double inva = 1.0 / gv.a, half = 0.5 * inva;
    const long nx = gv.n[0] + 1, ny = gv.n[1] + 1, nz = gv.n[2] + 1;
    out.assign((size_t)nx * ny * nz, 0.0);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < nz; ++k)
        for (long j = 0; j < ny; ++j)
            for (long i = 0; i < nx; ++i) {
                const vec r((2 * (int)i + (c == 0)) * half, (2 * (int)j + (c == 1)) * half, (2 * (int)k + (c == 2)) * half);
                out[((size_t)k * ny + j) * nx + i] = f(r);
            }
}

static const component e_comp[3] = {Ex, Ey, Ez};

structure::structure(const grid_volume &gv_, material_function &epsf, const boundary_region &br) : gv(gv_), pml_thickness(br.thickness) {
    if (gv.dim != D3) throw std::runtime_error("meep_compat: only 3-d grids (every junction conf)");
    for (int c = 0; c < 3; ++c) for_each_yee(gv, c, eps[c], [&](const vec &r) { return epsf.chi1p1(E_stuff, r); });
}
structure::~structure() {}

void structure::add_susceptibility(material_function &sigf, field_type ft, const susceptibility &s) {
    if (ft != E_stuff) throw std::runtime_error("meep_compat: only E_stuff susceptibilities");
    const lorentzian_susceptibility *ls = dynamic_cast<const lorentzian_susceptibility *>(&s);
    if (!ls) throw std::runtime_error("meep_compat: only lorentzian_susceptibility");
    sus.push_back(sus_rec());
    sus_rec &u = sus.back();
    u.omega_0 = ls->omega_0; u.gamma = ls->gamma; u.drude = ls->no_omega_0_denominator;
    for (int c = 0; c < 3; ++c)
        for_each_yee(gv, c, u.sigma[c], [&](const vec &r) {
            double row[3];
            sigf.sigma_row(e_comp[c], row, r);
            return row[component_index(e_comp[c])];
        });
}

// meep::fields(structure*): complex fields (two field sets), Courant 0.5, PML reflection 1e-15.  The per-point eps and
// sigma values collapse to a table of distinct materials (one byte per Yee point on the device).
fields::fields(structure *s) : dt(0), a(1), sim(0), strct(s), cached_step(-1), known_comp(-1) {
    a = s->gv.a;
    sj_grid g;
    memset(&g, 0, sizeof g);
    for (int d = 0; d < 3; ++d) g.n[d] = s->gv.n[d];
    g.a = s->gv.a; g.courant = 0.5; g.pml_thickness = s->pml_thickness; g.pml_R = 1e-15;
    const char *pr = getenv("SJ_PRECISION");
    g.precision = (pr && !strcmp(pr, "f32")) ? SJ_F32 : SJ_F64;
    g.n_sets = 2;
    g.device = getenv("SJ_DEVICE") ? atoi(getenv("SJ_DEVICE")) : -1;
    if (sj_create(&g, &sim)) throw std::runtime_error(std::string("sj_create: ") + sj_last_error(NULL));
    const size_t npt = s->eps[0].size(), nsus = s->sus.size();
    std::map<std::vector<double>, int> ids;
    std::vector<sj_material> mats;
    std::vector<uint8_t> idx[3];
    std::vector<double> key(1 + nsus);
    for (int c = 0; c < 3; ++c) {
        idx[c].resize(npt);
        for (size_t i = 0; i < npt; ++i) {
            key[0] = s->eps[c][i];
            for (size_t n = 0; n < nsus; ++n) key[1 + n] = s->sus[n].sigma[c][i];
            std::map<std::vector<double>, int>::iterator it = ids.find(key);
            if (it == ids.end()) {
                if (mats.size() >= 256) throw std::runtime_error("meep_compat: more than 256 distinct materials");
                sj_material M;
                memset(&M, 0, sizeof M);
                M.eps_inf = key[0];
                for (size_t n = 0; n < nsus; ++n)
                    if (key[1 + n] != 0.0) {
                        if (M.n_poles >= SJ_MAX_POLES) throw std::runtime_error("meep_compat: more than SJ_MAX_POLES poles at one point");
                        sj_pole &p = M.poles[M.n_poles++];
                        p.omega0 = s->sus[n].omega_0; p.gamma = s->sus[n].gamma; p.sigma = key[1 + n]; p.drude = s->sus[n].drude ? 1 : 0;
                    }
                it = ids.insert(std::make_pair(key, (int)mats.size())).first;
                mats.push_back(M);
            }
            idx[c][i] = (uint8_t)it->second;
        }
    }
    ck(sj_set_materials(sim, (int)mats.size(), mats.data(), idx[0].data(), idx[1].data(), idx[2].data()), sim, "sj_set_materials");
    dt = sj_dt(sim);
}
fields::~fields() {
    if (sim) sj_destroy(sim);
    for (size_t i = 0; i < srcs.size(); ++i) delete srcs[i];
}

static void dipole_trampoline(void *ctx, double time, double *out2) {
    const std::complex<double> d = static_cast<const src_time *>(ctx)->dipole(time);
    out2[0] = d.real(); out2[1] = d.imag();
}

void fields::add_volume_source(component c, const src_time &src, const volume &where, std::complex<double> amp) {
    if (!is_electric(c)) throw std::runtime_error("meep_compat: magnetic-current sources are not implemented");
    src_time *mine = src.clone();
    srcs.push_back(mine);
    const vec p1 = where.get_min_corner(), p2 = where.get_max_corner();
    const double lo[3] = {std::min(p1.x(), p2.x()), std::min(p1.y(), p2.y()), std::min(p1.z(), p2.z())};
    const double hi[3] = {std::max(p1.x(), p2.x()), std::max(p1.y(), p2.y()), std::max(p1.z(), p2.z())};
    ck(sj_add_custom_source(sim, component_index(c), lo, hi, amp.real(), amp.imag(), dipole_trampoline, mine, mine->last_time(),
                            mine->is_integrated ? 1 : 0, NULL), sim, "sj_add_custom_source");
}
void fields::add_point_source(component c, const src_time &src, const vec &p, std::complex<double> amp) {
    add_volume_source(c, src, volume(p, p), amp);
}
double fields::last_source_time() {
    double t = 0;
    for (size_t i = 0; i < srcs.size(); ++i) t = std::max(t, srcs[i]->last_time());
    return t;
}
void fields::set_output_directory(const char *dir) { outdir = dir ? dir : ""; }
volume fields::total_volume() const { return strct->gv.surroundings(); }
// fields.output_hdf5(meep::Dielectric, total_volume) (disp.cpp:696) -> <outdir>/eps-000000.00.h5; fields.output_hdf5(meep::Ex,
// vol.surroundings(), file) (disp.cpp:735) -> <outdir>/<file>.h5 with "ex.r" and "ex.i" (host/sj_dumps.hpp)
void fields::output_hdf5(component c, const volume &, h5file *file) {
    const int *n = strct->gv.n;
    const std::string dir = outdir.empty() ? std::string(".") : outdir;
    if (c == Dielectric) {
        if (sj_dump::write_eps(dir, strct->eps, n)) throw std::runtime_error("meep_compat: cannot write eps-000000.00.h5");
        return;
    }
    if (!is_electric(c)) throw std::runtime_error("meep_compat: output_hdf5 of Dielectric and E components only");
    static const char *nm[3] = {"ex", "ey", "ez"};
    const std::string name = file ? file->name : std::string(nm[component_index(c)]);
    if (sj_dump::write_field(dir + "/" + name + ".h5", sim, component_index(c), 2, n)) throw std::runtime_error("meep_compat: cannot write " + name + ".h5");
}
h5file *fields::open_h5file(const char *name) { return new h5file(name ? name : ""); }

std::complex<double> fields::get_field(component c, const vec &loc) const {
    const int comp = is_electric(c) ? component_index(c) : is_magnetic(c) ? 3 + component_index(c) : -1;
    if (comp < 0) throw std::runtime_error("meep_compat: get_field of E and H components only");
    const long long step = sj_steps_done(sim);
    if (comp != known_comp) { known_xyz.clear(); cached.clear(); known_comp = comp; cached_step = -1; }
    const size_t n = known_xyz.size() / 3;
    size_t at = n;
    for (size_t i = 0; i < n; ++i)
        if (known_xyz[3 * i] == loc.x() && known_xyz[3 * i + 1] == loc.y() && known_xyz[3 * i + 2] == loc.z()) { at = i; break; }
    if (at == n) {                       // a new point: remember it and fetch everything known so far
        known_xyz.push_back(loc.x()); known_xyz.push_back(loc.y()); known_xyz.push_back(loc.z());
        cached_step = -1;
    }
    if (cached_step != step) {
        cached.assign(known_xyz.size() / 3 * 2, 0.0);
        ck(sj_sample_at(sim, comp, (int)(known_xyz.size() / 3), known_xyz.data(), cached.data()), sim, "sj_sample_at");
        cached_step = step;
    }
    return std::complex<double>(cached[2 * at], cached[2 * at + 1]);
}
double fields::time() const { return sj_steps_done(sim) * dt; }
void fields::step() { ck(sj_run(sim, 1, 1 << 30), sim, "sj_run"); }

}  // namespace meep

#include "H5Cpp.h"
namespace H5 {
const PredType PredType::NATIVE_DOUBLE(1);
const PredType PredType::NATIVE_FLOAT(2);
const PredType PredType::NATIVE_HSIZE(3);
}
