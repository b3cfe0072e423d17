/*
 * host/meep_compat/meep.hpp -- the drop-in seam taken literally: the slice of meep's public C++ API that sim_juncs
 * touches (src/disp.cpp, src/disp.hpp, src/main.cpp), implemented on libsimjuncs_b200's C ABI
 * (include/sim_juncs_b200.h) by host/meep_compat/meep_compat.cpp.  With `-Ihost/meep_compat` in place of meep's include
 * directory and `-lsimjuncs_b200` in place of `-lmeep`, the reference's sources compile and link UNMODIFIED
 * (host/Makefile target sim_geom_meep) and its `sim_geom` steps on the B200 engine: params.conf, the .geom scene
 * language, cgs_material_function, gaussian_src_time_phase, the run loop and save_field_times stay the reference's code.
 *   meep::structure(gv, eps, pml) / add_susceptibility  -> the material_function virtuals are called at every Yee point
 *                                                          on the host (as meep does), distinct (eps, sigma...) tuples
 *                                                          become the material table of sj_set_materials
 *   meep::fields::add_volume_source(c, src, vol, amp)   -> sj_add_custom_source with src.clone()->dipole as the waveform
 *   meep::fields::step() / time() / dt                   -> sj_run(sim, 1, ...) / sj_steps_done * sj_dt / sj_dt
 *   meep::fields::get_field(c, loc)                      -> sj_sample_at, batched over the points seen so far
 *   meep::fields::last_source_time()                     -> max of src_time::last_time()
 * Every declaration names the reference call site that needs it.  Enumerator values and default arguments follow meep's
 * public header as recalled; no meep source is used.  (The faster route -- GPU rasterization of the CSG trees instead
 * of point-by-point virtual calls -- needs a ten-line change in bound_geom and is shown in INTEGRATION.md.)
 */
#ifndef SJ_MEEP_COMPAT_MEEP_HPP
#define SJ_MEEP_COMPAT_MEEP_HPP

#include <complex>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

struct sj_sim;

namespace meep {

extern int verbosity;                                  // main.cpp:33

// meep/vec.hpp enum order; disp.cpp:320-335 (source_info), :724 (get_field(meep::Ex, ...)), :696 (Dielectric)
enum component { Ex = 0, Ey, Er, Ep, Ez, Hx, Hy, Hr, Hp, Hz, Dx, Dy, Dr, Dp, Dz, Bx, By, Br, Bp, Bz, Dielectric, Permeability, NO_COMPONENT };
enum field_type { E_stuff = 0, H_stuff = 1, D_stuff = 2, B_stuff = 3, PE_stuff = 4, PH_stuff = 5, WE_stuff = 6, WH_stuff = 7 };
enum ndim { D1 = 0, D2, D3, Dcyl };

// disp.cpp:305: sigrow[meep::component_index(c)]
inline int component_index(component c) {
    switch (c) {
        case Ex: case Hx: case Dx: case Bx: case Er: case Hr: case Dr: case Br: return 0;
        case Ey: case Hy: case Dy: case By: case Ep: case Hp: case Dp: case Bp: return 1;
        case Ez: case Hz: case Dz: case Bz: return 2;
        default: return 3;
    }
}
inline bool is_electric(component c) { return c < Hx; }
inline bool is_magnetic(component c) { return c >= Hx && c < Dx; }

// disp.cpp:266 (r.x(), r.y(), r.z()), :463 (vec(x, y, z)), disp.hpp:146
class vec {
public:
    vec() : dim(D3) { t[0] = t[1] = t[2] = 0; }
    vec(double zz) : dim(D1) { t[0] = t[1] = 0; t[2] = zz; }
    vec(double xx, double yy) : dim(D2) { t[0] = xx; t[1] = yy; t[2] = 0; }
    vec(double xx, double yy, double zz) : dim(D3) { t[0] = xx; t[1] = yy; t[2] = zz; }
    double x() const { return t[0]; }
    double y() const { return t[1]; }
    double z() const { return t[2]; }
    ndim dim;
private:
    double t[3];
};

// disp.cpp:603 (meep::volume source_vol(vec, vec))
class volume {
public:
    volume() {}
    volume(const vec &a, const vec &b) : lo(a), hi(b) {}
    vec get_min_corner() const { return lo; }
    vec get_max_corner() const { return hi; }
private:
    vec lo, hi;
};

// disp.hpp:175 (member), disp.cpp:505-509 (vol1d/2d/3d), :735 (vol.surroundings())
class grid_volume {
public:
    grid_volume() : dim(D3), a(1) { n[0] = n[1] = n[2] = 0; }
    volume surroundings() const { return volume(vec(0, 0, 0), vec(n[0] / a, n[1] / a, n[2] / a)); }
    ndim dim;
    int n[3];
    double a;
};
grid_volume vol1d(double zsize, double a);
grid_volume vol2d(double xsize, double ysize, double a);
grid_volume vol3d(double xsize, double ysize, double zsize, double a);

// disp.hpp:55-85: the virtuals cgs_material_function overrides.  has_mu / has_conductivity / has_chi2 / has_chi3
// default to false in meep, so the mu / conductivity / chi2 / chi3 overrides are never called.
class material_function {
public:
    material_function() {}
    virtual ~material_function() {}
    virtual void set_volume(const volume &) {}
    virtual void unset_volume() {}
    virtual double chi1p1(field_type, const vec &) { return 1.0; }
    virtual double eps(const vec &) { return 1.0; }
    virtual bool has_mu() { return false; }
    virtual double mu(const vec &) { return 1.0; }
    virtual bool has_conductivity(component) { return false; }
    virtual double conductivity(component, const vec &) { return 0.0; }
    virtual void sigma_row(component c, double sigrow[3], const vec &) { sigrow[0] = sigrow[1] = sigrow[2] = 0.0; (void)c; }
    virtual bool has_chi3(component) { return false; }
    virtual double chi3(component, const vec &) { return 0.0; }
    virtual bool has_chi2(component) { return false; }
    virtual double chi2(component, const vec &) { return 0.0; }
};

// disp.hpp:112-133: base of gaussian_src_time_phase.  is_integrated is a public data member of meep's src_time,
// set to true by its constructor; the reference never touches it.
class src_time {
public:
    src_time() : is_integrated(true) {}
    src_time(const src_time &o) : is_integrated(o.is_integrated) {}
    virtual ~src_time() {}
    virtual std::complex<double> dipole(double) const { return 0; }
    virtual double last_time() const { return 0.0; }
    virtual src_time *clone() const { return new src_time(*this); }
    virtual bool is_equal(const src_time &) const { return false; }
    virtual std::complex<double> frequency() const { return 0.0; }
    virtual void set_frequency(std::complex<double>) {}
    bool is_integrated;
};

// disp.cpp:618: meep::continuous_src_time src(frequency, width, start_time, end_time)
class continuous_src_time : public src_time {
public:
    continuous_src_time(std::complex<double> f, double w = 0.0, double st = 0.0, double et = 1.0 / 0.0, double s = 3.0)
        : freq(f), width(w), start_time(st), end_time(et), slowness(s) {}
    virtual std::complex<double> dipole(double time) const;
    virtual double last_time() const { return end_time; }
    virtual src_time *clone() const { return new continuous_src_time(*this); }
    virtual std::complex<double> frequency() const { return freq; }
    virtual void set_frequency(std::complex<double> f) { freq = f; }
private:
    std::complex<double> freq;
    double width, start_time, end_time, slowness;
};

// disp.cpp:527: meep::pml(s.pml_thickness)
class boundary_region {
public:
    boundary_region(double t = 0.0) : thickness(t) {}
    double thickness;
};
boundary_region pml(double thickness);

// disp.cpp:542: meep::lorentzian_susceptibility suscept(omega_0, gamma, !use_denom)
class susceptibility {
public:
    virtual ~susceptibility() {}
};
class lorentzian_susceptibility : public susceptibility {
public:
    lorentzian_susceptibility(double omega_0, double gamma, bool no_omega_0_denominator = false)
        : omega_0(omega_0), gamma(gamma), no_omega_0_denominator(no_omega_0_denominator) {}
    double omega_0, gamma;
    bool no_omega_0_denominator;
};

// disp.cpp:527 (new meep::structure(vol, inf_eps_func, pml)), :546 (add_susceptibility), :648 (delete)
class structure {
public:
    structure(const grid_volume &gv, material_function &eps, const boundary_region &br = boundary_region());
    ~structure();
    void add_susceptibility(material_function &sigma, field_type ft, const susceptibility &sus);
    grid_volume gv;
    double pml_thickness;
    std::vector<double> eps[3];                                  // eps_inf at the Yee points of Ex, Ey, Ez
    struct sus_rec { double omega_0, gamma; bool drude; std::vector<double> sigma[3]; };
    std::vector<sus_rec> sus;
};

class h5file {                                                   // disp.cpp:734, :737
public:
    h5file(const std::string &name) : name(name) {}
    virtual ~h5file() {}
    std::string name;
};

// disp.hpp:177 (member), disp.cpp:560, 614, 619, 625, 659, 691, 696, 704, 724, 733-735, 740
class fields {
public:
    fields(structure *s);
    ~fields();
    void add_point_source(component c, const src_time &src, const vec &p, std::complex<double> amp = 1.0);
    void add_volume_source(component c, const src_time &src, const volume &where, std::complex<double> amp = 1.0);
    double last_source_time();
    void set_output_directory(const char *dir);
    volume total_volume() const;
    void output_hdf5(component c, const volume &where, h5file *file = 0);
    h5file *open_h5file(const char *name);
    std::complex<double> get_field(component c, const vec &loc) const;
    double time() const;
    void step();
    double dt;
    double a;
private:
    fields(const fields &);
    fields &operator=(const fields &);
    sj_sim *sim;
    structure *strct;
    std::vector<src_time *> srcs;
    std::string outdir;
    // get_field cache: the points asked for so far (the reference asks for the same monitor list at every save),
    // fetched from the device in one sj_sample_at call per save
    mutable std::vector<double> known_xyz, cached;
    mutable long long cached_step;
    mutable int known_comp;
};

class initialize {                                               // main.cpp:20
public:
    initialize(int &, char **&) {}
};

}  // namespace meep

#endif
