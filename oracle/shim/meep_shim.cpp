/*
 * oracle/shim/meep_shim.cpp -- TEST INFRASTRUCTURE.  Implements the meep API slice declared in oracle/shim/meep.hpp
 * on top of the CPU oracle (oracle/fdtd_oracle.c) and the HDF5 recorder declared in oracle/shim/H5Cpp.h.
 * See meep.hpp for what this is for.  Environment knobs (read by the shim, not by the reference):
 *   SJ_SHIM_DUMP=<dir>   write eps_{x,y,z}.f64 and sigma_<n>_{x,y,z}.f64 (what the reference's material functions
 *                        returned at every Yee point) and sources.txt into <dir>
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "meep.hpp"
#include "H5Cpp.h"

extern "C" {
orc_sim *orc_create(int nx, int ny, int nz, double a, double courant, double pml_thickness, double pml_R, int nsets);
void orc_destroy(orc_sim *s);
int orc_set_eps_pointwise(orc_sim *s, const double *ex, const double *ey, const double *ez);
int orc_add_sus_pointwise(orc_sim *s, double omega0, double gamma, int drude, const double *sx, const double *sy, const double *sz);
int orc_add_callback_source(orc_sim *s, int comp, const double *lo, const double *hi, double amp_re, double amp_im,
                            int integrated, void (*fn)(void *, double, double *), void *ctx);
void orc_get_field_at(const orc_sim *s, int comp, const double *xyz, double *out2);
void orc_step(orc_sim *s);
double orc_dt(const orc_sim *s);
long orc_time_steps(const orc_sim *s);
}

namespace meep {

int verbosity = 1;

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

// meep continuous_src_time::dipole [meep-recall]: same restatement as fdtd_oracle.c src_dipole, kind 1
std::complex<double> continuous_src_time::dipole(double time) const {
    float rtime = float(time);
    if (rtime < start_time || rtime > end_time) return 0.0;
    const double omega = 2 * M_PI * real(freq);
    const std::complex<double> amp = 1.0 / std::complex<double>(0, -omega);
    const std::complex<double> osc = std::complex<double>(cos(-omega * time), sin(-omega * time)) * amp;
    if (width == 0.0) return osc;
    const double ts = (time - start_time) / width - slowness;
    const double te = (end_time - time) / width - slowness;
    return osc * (1.0 + tanh(ts)) * (1.0 + tanh(te)) * 0.25;
}

static void dump_array(const char *name, const std::vector<double> &v) {
    const char *dir = getenv("SJ_SHIM_DUMP");
    if (!dir) return;
    std::string p = std::string(dir) + "/" + name;
    FILE *fp = fopen(p.c_str(), "wb");
    if (!fp) return;
    fwrite(v.data(), sizeof(double), v.size(), fp);
    fclose(fp);
}

// Yee point of E component c at array index (i, j, k): half-pixel index m = 2 i + (d == c), coordinate
// m * (0.5 / a) -- the same expression as oracle/csg_oracle.c and sim_juncs_b200/csrc/sj_raster.cu
template <typename F>
static void for_each_yee(const grid_volume &gv, int c, std::vector<double> &out, F f) {
    const double inva = 1.0 / gv.a, half = 0.5 * inva;
    const size_t nx = gv.n[0] + 1, ny = gv.n[1] + 1, nz = gv.n[2] + 1;
    out.assign(nx * ny * nz, 0.0);
    for (size_t k = 0; k < nz; ++k)
        for (size_t j = 0; j < ny; ++j)
            for (size_t i = 0; i < nx; ++i) {
                const vec r((2 * (int)i + (c == 0)) * half, (2 * (int)j + (c == 1)) * half, (2 * (int)k + (c == 2)) * half);
                out[(k * ny + j) * nx + i] = f(r);
            }
}

static const component e_comp[3] = {Ex, Ey, Ez};

// meep structure ctor -> structure_chunk::set_chi1inv(c, eps, ...) with use_anisotropic_averaging = false:
// chi1inv = 1 / eps.chi1p1(E_stuff, r) at each component's own Yee point
structure::structure(const grid_volume &gv_, material_function &epsf, const boundary_region &br) : gv(gv_), pml_thickness(br.thickness) {
    if (gv.dim != D3) throw std::runtime_error("oracle shim: only 3-d grids");
    for (int c = 0; c < 3; ++c) {
        for_each_yee(gv, c, eps[c], [&](const vec &r) { return epsf.chi1p1(E_stuff, r); });
        dump_array(c == 0 ? "eps_x.f64" : c == 1 ? "eps_y.f64" : "eps_z.f64", eps[c]);
    }
}
structure::~structure() {}

// meep structure_chunk::add_susceptibility: sigma[c][component_direction(c)] = sigma_row(c, row, r)[component_index(c)]
void structure::add_susceptibility(material_function &sigf, field_type ft, const susceptibility &s) {
    if (ft != E_stuff) throw std::runtime_error("oracle shim: only E_stuff susceptibilities");
    const lorentzian_susceptibility *ls = dynamic_cast<const lorentzian_susceptibility *>(&s);
    if (!ls) throw std::runtime_error("oracle shim: only lorentzian_susceptibility");
    sus.push_back(sus_rec());
    sus_rec &u = sus.back();
    u.omega_0 = ls->omega_0; u.gamma = ls->gamma; u.drude = ls->no_omega_0_denominator;
    for (int c = 0; c < 3; ++c) {
        for_each_yee(gv, c, u.sigma[c], [&](const vec &r) {
            double row[3];
            sigf.sigma_row(e_comp[c], row, r);
            return row[component_index(e_comp[c])];
        });
        char nm[64];
        snprintf(nm, sizeof nm, "sigma_%zu_%c.f64", sus.size() - 1, "xyz"[c]);
        dump_array(nm, u.sigma[c]);
    }
}

// meep fields(structure*): complex fields unless use_real_fields() is called (the reference never does), Courant 0.5,
// PML reflection 1e-15
fields::fields(structure *s) : dt(0), a(1), sim(0), strct(s) {
    a = s->gv.a;
    sim = orc_create(s->gv.n[0], s->gv.n[1], s->gv.n[2], s->gv.a, 0.5, s->pml_thickness, 1e-15, 2);
    orc_set_eps_pointwise(sim, s->eps[0].data(), s->eps[1].data(), s->eps[2].data());
    for (size_t n = 0; n < s->sus.size(); ++n)
        orc_add_sus_pointwise(sim, s->sus[n].omega_0, s->sus[n].gamma, s->sus[n].drude, s->sus[n].sigma[0].data(),
                              s->sus[n].sigma[1].data(), s->sus[n].sigma[2].data());
    dt = orc_dt(sim);
}
fields::~fields() {
    if (sim) orc_destroy(sim);
    for (size_t i = 0; i < srcs.size(); ++i) delete srcs[i];
}

static void dipole_trampoline(void *ctx, double time, double *out2) {
    const std::complex<double> d = static_cast<const src_time *>(ctx)->dipole(time);
    out2[0] = d.real(); out2[1] = d.imag();
}

// meep fields::add_volume_source(c, src, where, amp): src is cloned; placement weights are the oracle's
void fields::add_volume_source(component c, const src_time &src, const volume &where, std::complex<double> amp) {
    if (!is_electric(c)) throw std::runtime_error("oracle shim: magnetic-current sources are not restated");
    src_time *mine = src.clone();
    srcs.push_back(mine);
    const vec p1 = where.get_min_corner(), p2 = where.get_max_corner();
    // meep::volume(vec, vec) stores the componentwise min and max corners
    const double lo[3] = {std::min(p1.x(), p2.x()), std::min(p1.y(), p2.y()), std::min(p1.z(), p2.z())};
    const double hi[3] = {std::max(p1.x(), p2.x()), std::max(p1.y(), p2.y()), std::max(p1.z(), p2.z())};
    const int rc = orc_add_callback_source(sim, component_index(c), lo, hi, amp.real(), amp.imag(), mine->is_integrated ? 1 : 0,
                                           dipole_trampoline, mine);
    if (rc) throw std::runtime_error("oracle shim: source placement failed");
    const char *dir = getenv("SJ_SHIM_DUMP");
    if (dir) {
        FILE *fp = fopen((std::string(dir) + "/sources.txt").c_str(), "a");
        if (fp) {
            fprintf(fp, "%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)c, lo[0], lo[1], lo[2], hi[0], hi[1],
                    hi[2], amp.real(), amp.imag(), mine->last_time());
            fclose(fp);
        }
    }
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
void fields::output_hdf5(component, const volume &, h5file *) {}       // eps-*.h5 / ex-*.h5 dumps: not recorded
h5file *fields::open_h5file(const char *name) { return new h5file(name ? name : ""); }
std::complex<double> fields::get_field(component c, const vec &loc) const {
    if (!is_electric(c)) throw std::runtime_error("oracle shim: get_field of E components only");
    const double xyz[3] = {loc.x(), loc.y(), loc.z()};
    double o[2];
    orc_get_field_at(sim, component_index(c), xyz, o);
    return std::complex<double>(o[0], o[1]);
}
double fields::time() const { return orc_time_steps(sim) * dt; }
void fields::step() { orc_step(sim); }

}  // namespace meep

// ---------------------------------------------------------------------------------------------------
// HDF5 recorder
// ---------------------------------------------------------------------------------------------------
namespace H5 {

const PredType PredType::NATIVE_DOUBLE(1);
const PredType PredType::NATIVE_FLOAT(2);
const PredType PredType::NATIVE_HSIZE(3);

static std::vector<shim_type> &type_table() {
    static std::vector<shim_type> t;
    if (t.empty()) {
        t.resize(4);
        t[1].kind = "f64"; t[1].size = 8;
        t[2].kind = "f32"; t[2].size = 4;
        t[3].kind = "u64"; t[3].size = 8;
    }
    return t;
}
shim_type &shim_type_of(hid_t id) { return type_table().at((size_t)id); }
hid_t shim_new_compound(size_t size) {
    std::vector<shim_type> &t = type_table();
    shim_type c;
    c.kind = "compound"; c.size = size;
    t.push_back(c);
    return (hid_t)t.size() - 1;
}

struct shim_entry { char what; std::string path; hid_t type; std::vector<hsize_t> dims; long long offset; size_t nbytes; };
struct shim_file { std::string name; std::vector<shim_entry> entries; std::vector<unsigned char> blob; bool open; };

static std::string type_json(hid_t id) {
    const shim_type &t = shim_type_of(id);
    char b[256];
    if (t.kind != "compound") { snprintf(b, sizeof b, "{\"kind\": \"%s\", \"size\": %zu}", t.kind.c_str(), t.size); return b; }
    snprintf(b, sizeof b, "{\"kind\": \"compound\", \"size\": %zu, \"members\": [", t.size);
    std::string s = b;
    for (size_t i = 0; i < t.members.size(); ++i) {
        snprintf(b, sizeof b, "%s{\"name\": \"%s\", \"offset\": %zu, \"type\": \"%s\"}", i ? ", " : "", t.members[i].name.c_str(),
                 t.members[i].offset, shim_type_of(t.members[i].type).kind.c_str());
        s += b;
    }
    return s + "]}";
}

void DataSet::write(const void *buf, const DataType &mem_type) {
    if (!f) return;
    shim_entry &e = f->entries[index];
    size_t n = shim_type_of(mem_type.getId()).size;
    for (size_t i = 0; i < e.dims.size(); ++i) n *= (size_t)e.dims[i];
    e.offset = (long long)f->blob.size();
    e.nbytes = n;
    const unsigned char *p = static_cast<const unsigned char *>(buf);
    f->blob.insert(f->blob.end(), p, p + n);
}
Group Group::createGroup(const char *name) {
    Group g;
    g.f = f; g.path = path + "/" + name;
    shim_entry e; e.what = 'G'; e.path = g.path; e.type = 0; e.offset = -1; e.nbytes = 0;
    f->entries.push_back(e);
    return g;
}
DataSet Group::createDataSet(const char *name, const DataType &type, const DataSpace &space) {
    // libhdf5 rejects a NULL name and its C++ layer throws; the reference does not catch it (an unnamed numeric in the
    // scene context, e.g. an expression statement, ends save_field_times this way)
    if (!name) throw std::runtime_error("H5::Group::createDataSet: NULL name");
    DataSet d;
    d.f = f; d.index = f->entries.size();
    shim_entry e; e.what = 'D'; e.path = path + "/" + name; e.type = type.getId(); e.dims = space.dims; e.offset = -1; e.nbytes = 0;
    f->entries.push_back(e);
    return d;
}
H5File::H5File(const char *name, unsigned) {
    f = new shim_file;
    f->name = name; f->open = true;
    path = "";
}
void H5File::close() {
    if (!f || !f->open) return;
    f->open = false;
    FILE *fj = fopen((f->name + ".manifest.json").c_str(), "w");
    FILE *fb = fopen((f->name + ".manifest.bin").c_str(), "wb");
    if (!fj || !fb) { fprintf(stderr, "oracle shim: cannot write %s.manifest.*\n", f->name.c_str()); if (fj) fclose(fj); if (fb) fclose(fb); return; }
    fprintf(fj, "{\"file\": \"%s\", \"entries\": [\n", f->name.c_str());
    for (size_t i = 0; i < f->entries.size(); ++i) {
        const shim_entry &e = f->entries[i];
        if (e.what == 'G') fprintf(fj, "  {\"what\": \"group\", \"path\": \"%s\"}", e.path.c_str());
        else {
            fprintf(fj, "  {\"what\": \"dataset\", \"path\": \"%s\", \"type\": %s, \"dims\": [", e.path.c_str(), type_json(e.type).c_str());
            for (size_t d = 0; d < e.dims.size(); ++d) fprintf(fj, "%s%llu", d ? ", " : "", e.dims[d]);
            fprintf(fj, "], \"offset\": %lld, \"nbytes\": %zu}", e.offset, e.nbytes);
        }
        fprintf(fj, "%s\n", i + 1 < f->entries.size() ? "," : "");
    }
    fprintf(fj, "]}\n");
    fwrite(f->blob.data(), 1, f->blob.size(), fb);
    fclose(fj); fclose(fb);
}
H5File::~H5File() { close(); delete f; f = 0; }

}  // namespace H5

herr_t H5Tinsert(hid_t compound, const char *name, size_t offset, hid_t member) {
    H5::shim_type &t = H5::shim_type_of(compound);
    if (t.kind != "compound") return -1;
    H5::shim_member m; m.name = name; m.offset = offset; m.type = member;
    t.members.push_back(m);
    return 0;
}
