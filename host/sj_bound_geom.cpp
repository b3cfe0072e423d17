// sj_bound_geom.cpp -- `bound_geom` on the C ABI.  Construction order, unit conversions and the run
// loop cadence are those of reference src/disp.cpp:482-749; every meep call is replaced by an sj_* call.
#include <algorithm>
#include "sj_host.hpp"
#include "sj_dumps.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// ---- evaluation context (reference src/disp.cpp:28-51) -------------------------------------------
static value h_um_to_l(context &c, cgs_func f, parse_ercode &er) {
    value none;
    if (f.n_args < 1) { er = E_LACK_TOKENS; return none; }
    if (f.args[0].type != VAL_NUM) { er = E_BAD_TOKEN; return none; }
    value scale = c.lookup("l_per_um");
    if (scale.type != VAL_NUM) { er = E_NOT_DEFINED; return none; }
    return make_val_num(scale.val.x * f.args[0].val.x);
}
static value h_fs_to_t(context &c, cgs_func f, parse_ercode &er) {
    value none;
    if (f.n_args < 1) { er = E_LACK_TOKENS; return none; }
    if (f.args[0].type != VAL_NUM) { er = E_BAD_TOKEN; return none; }
    value scale = c.lookup("l_per_um");
    if (scale.type != VAL_NUM) { er = E_NOT_DEFINED; return none; }
    return make_val_num(SJ_LIGHT_SPEED * scale.val.x * f.args[0].val.x);
}

context sj_context_from_settings(const parse_settings &a) {
    context con;
    con.emplace("pi", make_val_num(M_PI));
    con.emplace("pml_thickness", make_val_num(a.pml_thickness));
    con.emplace("sim_length", make_val_num(a.len));
    con.emplace("length", make_val_num(2 * a.pml_thickness + a.len));
    con.emplace("l_per_um", make_val_num(a.um_scale));
    value tmp = make_val_str(a.out_dir);
    con.emplace("out_dir", tmp);
    cleanup_val(&tmp);
    tmp = make_val_func("um_to_l", 1, &h_um_to_l);
    con.emplace("um_to_l", tmp);
    cleanup_val(&tmp);
    tmp = make_val_func("fs_to_t", 1, &h_fs_to_t);
    con.emplace("fs_to_t", tmp);
    cleanup_val(&tmp);
    if (a.user_opts) { line_buffer lb(a.user_opts, ';'); con.read_from_lines(lb); }
    return con;
}

// ---- source_info (reference src/disp.cpp:318-376) -------------------------------------------------
static bool num_field(value inst, const char *key, double *out) {
    value v = inst.val.c->lookup(key);
    if (v.type != VAL_NUM) return false;
    *out = v.val.x;
    return true;
}
sj_source_info::sj_source_info(value info) {
    type = SJ_SRC_GAUSSIAN; component = 0; wavelen = 0.7; width = 1; phase = 0; start_time = 0; end_time = 0; amplitude = 1.0;
    const bool gauss = is_type(info, "Gaussian_source"), contin = is_type(info, "CW_source");
    ok = gauss || contin;
    if (!ok) return;
    double x;
    if (num_field(info, "component", &x) && (int)x >= 1 && (int)x <= 5) component = (int)x;
    num_field(info, "wavelength", &wavelen);
    num_field(info, "amplitude", &amplitude);
    num_field(info, "start_time", &start_time);
    num_field(info, "end_time", &end_time);
    num_field(info, "slowness", &width);
    if (gauss) {
        num_field(info, "width", &width);
        num_field(info, "phase", &phase);
        double cutoff = 5;
        if (num_field(info, "cutoff", &x)) cutoff = 6;   // sic (disp.cpp:369-372)
        end_time = start_time + 2 * cutoff * width;
    }
    if (contin) type = SJ_SRC_CONTINUOUS;
}

// ---- parse_susceptibilities (reference src/disp.cpp:418-454) ----------------------------------------
std::vector<sj_pole_raw> sj_bound_geom::parse_susceptibilities(value val, int *er) {
    std::vector<sj_pole_raw> ret;
    if (val.type == VAL_LIST) {
        for (size_t i = 0; i < val.n_els; ++i) {
            value cur = val.val.l[i];
            if (cur.type != VAL_LIST || cur.n_els < 2) { if (er) *er = -1; return ret; }
            if (cur.n_els < 3 || cur.val.l[0].type != VAL_NUM || cur.val.l[1].type != VAL_NUM || cur.val.l[2].type != VAL_NUM) {
                if (er) *er = -2;
                return ret;
            }
            sj_pole_raw p = {cur.val.l[0].val.x, cur.val.l[1].val.x, cur.val.l[2].val.x, true};
            if (cur.n_els > 3) {
                if (cur.val.l[3].type == VAL_NUM) p.use_denom = (cur.val.l[3].val.x != 0.0);
                else if (cur.val.l[3].type == VAL_STR) {
                    const char *tok = cur.val.l[3].val.s;
                    if (!strcmp(tok, "drude") || !strcmp(tok, "true")) p.use_denom = false;
                    if (!strcmp(tok, "lorentz") || !strcmp(tok, "false")) p.use_denom = true;
                } else { if (er) *er = -2; return ret; }
            }
            ret.push_back(p);
        }
    }
    if (er) *er = 0;
    return ret;
}

// ---- constructor (reference src/disp.cpp:482-645) -----------------------------------------------
sj_bound_geom::sj_bound_geom(const parse_settings &s, parse_ercode *ercode, int precision, int p_n_sets, int integrated, int n_gpus)
    : problem(s.geom_fname, sj_context_from_settings(s), ercode), sim(NULL) {
    if (ercode && *ercode != E_SUCCESS) { printf("Scene parsing failed, exiting.\n"); exit(1); }
    um_scale = s.um_scale; post_source_t = s.post_source_t; save_span = s.save_span; n_sets = p_n_sets;
    ttot = 0; n_t_pts = 0; raster_ms = run_s = 0; dump_raw = s.dump_raw;
    if (s.n_dims != 3) { fprintf(stderr, "only dimensions = 3 is supported\n"); exit(1); }
    printf("using simulation side length %f, resolution %f\n", s.len, s.resolution);

    const double z_center = s.len / 2 + s.pml_thickness;
    sj_grid g;
    memset(&g, 0, sizeof g);
    const int n = (int)(2 * z_center * s.resolution + 0.5);          // meep::vol3d
    g.n[0] = g.n[1] = g.n[2] = n; n_cells[0] = n_cells[1] = n_cells[2] = n; g.a = s.resolution; g.courant = 0.5; g.pml_thickness = s.pml_thickness; g.pml_R = 1e-15;
    g.precision = precision; g.n_sets = n_sets; g.device = -1;
    n_gpus = std::max(1, std::min(n_gpus, n + 1));

    std::vector<composite_object *> roots = problem.get_roots();
    std::vector<sj_csg_node> nodes;
    std::vector<sj_region> regions(roots.size());
    for (size_t i = 0; i < roots.size(); ++i) {
        double thick = 1.0;
        if (roots[i]->has_metadata("make_2d") && roots[i]->fetch_metadata("make_2d").val.x != 0) {
            thick = SJ_THICK_SCALE / s.resolution;
            roots[i]->rescale(vec3(1.0, 1.0, thick));                 // disp.cpp:521-524
        }
        memset(&regions[i], 0, sizeof(sj_region));
        regions[i].root = sj_flatten_tree(roots[i], nodes);
        regions[i].eps = s.ambient_eps;                               // add_region(): default scale
        if (roots[i]->has_metadata("eps")) {
            value v = roots[i]->fetch_metadata("eps");
            if (v.type == VAL_NUM) regions[i].eps = v.val.x;
        }
        std::vector<sj_pole_raw> sus;
        int res = 0;
        if (roots[i]->has_metadata("susceptibilities")) sus = parse_susceptibilities(roots[i]->fetch_metadata("susceptibilities"), &res);
        for (size_t j = 0; j < sus.size() && j < SJ_MAX_POLES; ++j) {
            sj_pole &p = regions[i].poles[regions[i].n_poles++];
            p.omega0 = sus[j].omega_0 / s.um_scale; p.gamma = sus[j].gamma / s.um_scale;    // disp.cpp:539-541
            p.sigma = sus[j].sigma / thick; p.drude = !sus[j].use_denom;
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    // one sj_sim per z-slab: created and rasterized on its GPU (a slab rasterizes only its own planes)
    auto build = [&](const std::vector<int> &cut) {
        for (sj_sim *q : sims) sj_destroy(q);
        sims.assign(n_gpus, NULL);
        for (int r = 0; r < n_gpus; ++r) {
            sj_grid gr = g;
            if (n_gpus > 1) { gr.kz0 = cut[r]; gr.kz1 = cut[r + 1]; gr.device = getenv("SJ_ONE_DEVICE") ? 0 : r; }   // SJ_ONE_DEVICE: all slabs on GPU 0 (tests)
            if (sj_create(&gr, &sims[r])) { fprintf(stderr, "sj_create: %s\n", sj_last_error(sims[r])); exit(1); }
            if (sj_rasterize_smooth(sims[r], s.ambient_eps, (int)nodes.size(), nodes.data(), (int)regions.size(), regions.data(),
                                    (int)s.smooth_n, s.smooth_rad)) {
                fprintf(stderr, "sj_rasterize_smooth: %s\n", sj_last_error(sims[r])); exit(1);
            }
        }
    };
    std::vector<int> cut(n_gpus + 1);
    for (int r = 0; r <= n_gpus; ++r) cut[r] = (int)((long long)(n + 1) * r / n_gpus);
    build(cut);
    if (n_gpus > 1) {       // cut again on the measured bytes per plane (substrate planes carry polarisation traffic)
        std::vector<double> w(n + 1);
        for (int r = 0; r < n_gpus; ++r) sj_plane_costs(sims[r], &w[cut[r]]);
        double tot = 0, acc = 0; for (double x : w) tot += x;
        int target = 1;
        for (int k = 0; k <= n && target < n_gpus; ++k) {
            acc += w[k];
            while (target < n_gpus && acc >= tot * target / n_gpus) cut[target++] = k + 1;
        }
        for (int r = 1; r < n_gpus; ++r) cut[r] = std::max(cut[r], cut[r - 1] + 1);
        for (int r = n_gpus - 1; r > 0; --r) cut[r] = std::min(cut[r], cut[r + 1] - 1);
        build(cut);
        for (int r = 0; r + 1 < n_gpus; ++r)
            if (sj_connect_local(sims[r], sims[r + 1])) { fprintf(stderr, "sj_connect_local: %s\n", sj_last_error(sims[r])); exit(1); }
        printf("z-slabs:");
        for (int r = 0; r < n_gpus; ++r) printf(" [%d,%d)", cut[r], cut[r + 1]);
        printf("\n");
    }
    sim = sims[0];
    raster_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

    context &c = problem.get_context();
    const double c_by_a = SJ_LIGHT_SPEED * s.um_scale;
    for (size_t i = c.size(); i > 0; --i) {
        value inst = c.peek_val(i);
        sj_source_info info(inst);
        if (info.ok) {
            value reg = inst.val.c->lookup("region");
            if (!is_type(reg, "Box")) continue;
            parse_ercode e1 = E_SUCCESS, e2 = E_SUCCESS;
            value p1 = reg.val.c->lookup("pt_1").cast_to(VAL_3VEC, e1);
            value p2 = reg.val.c->lookup("pt_2").cast_to(VAL_3VEC, e2);
            if (e1 == E_SUCCESS && e2 == E_SUCCESS) {
                double lo[3], hi[3];
                for (int d = 0; d < 3; ++d) {
                    lo[d] = std::min(p1.val.v->el[d], p2.val.v->el[d]);
                    hi[d] = std::max(p1.val.v->el[d], p2.val.v->el[d]);
                }
                const double frequency = 1 / (info.wavelen * s.um_scale), width = info.width * c_by_a;
                const double t_start = info.start_time * c_by_a, t_end = info.end_time * c_by_a;
                if (info.type == SJ_SRC_GAUSSIAN && info.component <= 2) {
                    printf("Adding Gaussian envelope: f=%f, w=%f, t_0=%f, t_f=%f (meep units)\n", frequency, width, t_start, t_end);
                    for (sj_sim *q : sims)
                        if (sj_add_gaussian_source(q, info.component, lo, hi, info.amplitude, 0.0, frequency, width, info.phase, t_start,
                                                   t_end, integrated, NULL)) { fprintf(stderr, "source: %s\n", sj_last_error(q)); exit(1); }
                } else if (info.component <= 2) {
                    printf("Adding continuous wave: f=%f, w=%f, t_0=%f, t_f=%f (meep units)\n", frequency, width, t_start, t_end);
                    for (sj_sim *q : sims)
                        if (sj_add_cw_source(q, info.component, lo, hi, info.amplitude, 0.0, frequency, width, t_start, t_end, 3.0,
                                             integrated, NULL)) { fprintf(stderr, "source: %s\n", sj_last_error(q)); exit(1); }
                } else {
                    // same behaviour as the Python host (bound_geom.py): no silent change of the simulated problem
                    fprintf(stderr, "error: magnetic-current sources (component %d) are not implemented in the CUDA engine\n", (int)info.component);
                    exit(1);
                }
                sources.push_back(info);
                ttot = sj_last_source_time(sim) + post_source_t * SJ_LIGHT_SPEED * s.um_scale;
            }
            cleanup_val(&p1);
            cleanup_val(&p2);
        } else if (is_type(inst, "monitor")) {
            value vl = inst.val.c->lookup("locations");
            if (vl.type == VAL_LIST) {
                for (size_t q = 0; q < vl.n_els; ++q) {
                    parse_ercode er = E_SUCCESS;
                    value vc = vl.val.l[q].cast_to(VAL_3VEC, er);
                    if (er != E_SUCCESS) { cleanup_val(&vc); printf("warning: Invalid monitor location encountered!\n"); break; }
                    sj_vec3 loc = {vc.val.v->x(), vc.val.v->y(), vc.val.v->z()};
                    monitor_locs.push_back(loc);
                    cleanup_val(&vc);
                }
            }
            monitor_clusters.push_back(monitor_locs.size());
        }
    }
    for (sj_sim *q : sims)
        if (!monitor_locs.empty() && sj_add_monitors(q, SJ_EX, (int)monitor_locs.size(), &monitor_locs[0].x)) {
            fprintf(stderr, "monitors: %s\n", sj_last_error(q)); exit(1);
        }
    if (ercode) *ercode = E_SUCCESS;
}

sj_bound_geom::~sj_bound_geom() { for (sj_sim *q : sims) sj_destroy(q); }

// ---- run (reference src/disp.cpp:690-749) ---------------------------------------------------------
int sj_bound_geom::run(const char *fname_prefix) {
    printf("Set output directory to %s\n", fname_prefix);
    if (save_span == 0) save_span = 1;
    const double dt = sj_dt(sim);
    n_t_pts = (unsigned)((ttot + dt / 2) / dt);
    printf("starting simulations\n");
    auto t0 = std::chrono::steady_clock::now();
    int rc = 0;
    if (sims.size() > 1) {
        if (dump_raw) printf("warning: whole-grid dumps (eps-*.h5, ex-*.h5) are written by single-GPU runs only\n");
        rc = sj_run_group(sims.data(), (int)sims.size(), n_t_pts, save_span);
    } else if (fname_prefix && sj_dump::write_eps_from_sim(fname_prefix, sim, n_cells))      // fields.output_hdf5(Dielectric), disp.cpp:696
        printf("warning: could not write %s/eps-000000.00.h5\n", fname_prefix);
    if (sims.size() > 1) {
    } else if (dump_raw && fname_prefix) {
        // disp.cpp:709-737: at every save point the whole Ex field goes to ex-<time>.h5 before the step
        const int n_digits_a = (int)(ceil(log(ttot) / log(10)));
        const int n_digits_b = (int)ceil(-log((double)dt) / log(10.0)) + 1;
        char h5_fname[BUF_SIZE];
        strcpy(h5_fname, "ex-");
        for (unsigned done = 0; done < n_t_pts && !rc; done += save_span) {
            make_dec_str(h5_fname + 3, BUF_SIZE - 3, done * dt, n_digits_a, n_digits_b);
            if (sj_dump::write_field(std::string(fname_prefix) + "/" + h5_fname + ".h5", sim, SJ_EX, n_sets, n_cells))
                printf("warning: could not write %s\n", h5_fname);
            rc = sj_run(sim, std::min(save_span, n_t_pts - done), save_span);
            if (rc == SJ_ERR_DIVERGED) rc = 0;
        }
    } else rc = sj_run(sim, n_t_pts, save_span);
    for (sj_sim *q : sims) { const int r2 = sj_sync(q); if (!rc) rc = r2; }
    if (rc == SJ_ERR_DIVERGED) printf("divergence in run (%s)\n", sj_last_error(sim));
    else if (rc) { printf("error in run: %s\n", sj_last_error(sim)); return rc; }
    run_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const size_t n_locs = monitor_locs.size(), ns = (size_t)sj_n_samples(sim);
    std::vector<double> buf(std::max<size_t>(ns * n_locs * n_sets, 1)), part(buf.size());
    sj_read_monitors(sim, buf.data());
    for (size_t r = 1; r < sims.size(); ++r) {          // every monitor is evaluated by the slab that owns it, zeros elsewhere
        sj_read_monitors(sims[r], part.data());
        for (size_t q = 0; q < buf.size(); ++q) buf[q] += part[q];
    }
    field_times.assign(n_locs, std::vector<std::complex<double> >(ns));
    for (size_t i = 0; i < ns; ++i)
        for (size_t j = 0; j < n_locs; ++j) {
            const double *q = &buf[(i * n_locs + j) * n_sets];
            field_times[j][i] = std::complex<double>(q[0], n_sets > 1 ? q[1] : 0.0);
        }
    printf("Simulations completed\n");
    return 0;
}

// ---- save_field_times: same datasets as src/disp.cpp:758-923, as field_samples.npz ------------------
// (libhdf5 is not available in this image; keys are the HDF5 paths, compound types become 2-D arrays)
static uint32_t crc32_of(const unsigned char *p, size_t n) {
    static uint32_t tab[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; tab[i] = c; } init = true; }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
struct NpzWriter {
    FILE *fp; std::vector<std::string> names; std::vector<uint32_t> crcs, sizes, offs;
    explicit NpzWriter(const char *path) { fp = fopen(path, "wb"); }
    static void le16(std::string &s, unsigned v) { s.push_back((char)(v & 255)); s.push_back((char)((v >> 8) & 255)); }
    static void le32(std::string &s, uint32_t v) { for (int i = 0; i < 4; ++i) s.push_back((char)((v >> (8 * i)) & 255)); }
    void add(const std::string &key, const double *data, size_t rows, size_t cols) {
        if (!fp) return;
        char shape[64];
        if (cols) snprintf(shape, sizeof shape, "(%zu, %zu)", rows, cols); else snprintf(shape, sizeof shape, "(%zu,)", rows);
        std::string hdr = std::string("{'descr': '<f8', 'fortran_order': False, 'shape': ") + shape + ", }";
        while ((10 + hdr.size() + 1) % 64) hdr.push_back(' ');
        hdr.push_back('\n');
        std::string body("\x93NUMPY\x01\x00", 8);
        le16(body, (unsigned)hdr.size());
        body += hdr;
        const size_t nbytes = rows * (cols ? cols : 1) * sizeof(double);
        body.append((const char *)data, nbytes);
        const std::string fname = key + ".npy";
        const uint32_t crc = crc32_of((const unsigned char *)body.data(), body.size());
        std::string lh; le32(lh, 0x04034b50u); le16(lh, 20); le16(lh, 0); le16(lh, 0); le16(lh, 0); le16(lh, 0);
        le32(lh, crc); le32(lh, (uint32_t)body.size()); le32(lh, (uint32_t)body.size()); le16(lh, (unsigned)fname.size()); le16(lh, 0);
        offs.push_back((uint32_t)ftell(fp)); names.push_back(fname); crcs.push_back(crc); sizes.push_back((uint32_t)body.size());
        fwrite(lh.data(), 1, lh.size(), fp); fwrite(fname.data(), 1, fname.size(), fp); fwrite(body.data(), 1, body.size(), fp);
    }
    void close() {
        if (!fp) return;
        const uint32_t cd_off = (uint32_t)ftell(fp);
        uint32_t cd_size = 0;
        for (size_t i = 0; i < names.size(); ++i) {
            std::string ch; le32(ch, 0x02014b50u); le16(ch, 20); le16(ch, 20); le16(ch, 0); le16(ch, 0); le16(ch, 0); le16(ch, 0);
            le32(ch, crcs[i]); le32(ch, sizes[i]); le32(ch, sizes[i]); le16(ch, (unsigned)names[i].size()); le16(ch, 0); le16(ch, 0);
            le16(ch, 0); le16(ch, 0); le32(ch, 0); le32(ch, offs[i]);
            fwrite(ch.data(), 1, ch.size(), fp); fwrite(names[i].data(), 1, names[i].size(), fp);
            cd_size += (uint32_t)(ch.size() + names[i].size());
        }
        std::string e; le32(e, 0x06054b50u); le16(e, 0); le16(e, 0); le16(e, (unsigned)names.size()); le16(e, (unsigned)names.size());
        le32(e, cd_size); le32(e, cd_off); le16(e, 0);
        fwrite(e.data(), 1, e.size(), fp);
        fclose(fp); fp = NULL;
    }
};

int sj_bound_geom::save_field_times(const char *fname_prefix) {
    char path[1024];
    snprintf(path, sizeof path, "%s/field_samples.h5", fname_prefix);
    printf("saving field output to %s\n", path);
    {   // field_samples.h5: the reference's groups, names and compound types (disp.cpp:758-923), own encoder
        sj_h5::Writer h;
        typedef std::vector<std::pair<std::string, unsigned> > members;
        members cplx, loc, src;
        cplx.push_back(std::make_pair("Re", 0u)); cplx.push_back(std::make_pair("Im", 8u));
        loc.push_back(std::make_pair("x", 0u)); loc.push_back(std::make_pair("y", 8u)); loc.push_back(std::make_pair("z", 16u));
        const char *sn[6] = {"wavelen", "width", "phase", "start_time", "end_time", "amplitude"};
        for (int q = 0; q < 6; ++q) src.push_back(std::make_pair(std::string(sn[q]), 8u + 8u * q));   // HOFFSET in class source_info
        const sj_h5::bytes t_cplx = sj_h5::compound_type(cplx, 16), t_loc = sj_h5::compound_type(loc, 24), t_src = sj_h5::compound_type(src, 56);
        const size_t n_locs = monitor_locs.size();
        const double ttot_fs = meep_time_to_fs(ttot);
        const double tb[3] = {0.0, ttot_fs, n_t_pts ? ttot_fs * save_span / n_t_pts : 0.0};
        h.dataset_f64("info/time_bounds", tb, 3);
        uint64_t u = monitor_clusters.size(); h.dataset_u64("info/n_clusters", &u, 1);
        u = n_t_pts / save_span; h.dataset_u64("info/n_time_points", &u, 1);
        std::vector<double> sb(sources.size() * 7, 0.0);
        for (size_t q = 0; q < sources.size(); ++q) {
            const sj_source_info &g = sources[q];
            const double row[6] = {g.wavelen, g.width, g.phase, g.start_time, g.end_time, g.amplitude};
            memcpy(&sb[7 * q + 1], row, sizeof row);
        }
        h.dataset("info/sources", t_src, sb.data(), sources.size(), 56);
        h.group("info/cgs_params");
        context &cx = problem.get_context();
        for (size_t q = cx.size(); q > 0; --q) {
            name_val_pair nv = cx.peek(q);
            if (nv.get_val().type == VAL_NUM && nv.get_name() && nv.get_name()[0]) {
                const double x = nv.get_val().val.x;
                h.dataset_f64(std::string("info/cgs_params/") + nv.get_name(), &x, 1);
            }
        }
        // `frequency` (data_utils.cpp:370-427 as called at disp.cpp:806) on the device: [n_locs][n_freq]{re, im}
        int32_t n_freq = 0;
        sj_read_spectra(sim, 0, n_sets > 1 ? 1 : -1, &n_freq, NULL);
        std::vector<double> spec(std::max<size_t>((size_t)n_locs * n_freq * 2, 1));
        if (n_freq && n_locs && sj_read_spectra(sim, 0, n_sets > 1 ? 1 : -1, &n_freq, spec.data())) { printf("%s\n", sj_last_error(sim)); return -1; }
        for (size_t r = 1; r < sims.size() && n_freq && n_locs; ++r) {      // the transform is linear: sum over the slabs
            std::vector<double> part(spec.size());
            if (sj_read_spectra(sims[r], 0, n_sets > 1 ? 1 : -1, &n_freq, part.data())) { printf("%s\n", sj_last_error(sims[r])); return -1; }
            for (size_t q = 0; q < spec.size(); ++q) spec[q] += part[q];
        }
        const size_t ngd = (size_t)(log((double)std::max<size_t>(monitor_clusters.size(), 1)) / log(10.0)) + 1;
        const size_t npd = (size_t)(log((double)std::max<size_t>(n_locs, 1)) / log(10.0)) + 1;
        size_t i = 0, off = 0;
        for (size_t j = 0; j < monitor_clusters.size() + 1; ++j) {      // sic: one empty trailing cluster (disp.cpp:879)
            const size_t max_i = j >= monitor_clusters.size() ? n_locs : monitor_clusters[j];
            char cname[64], pname[64];
            strcpy(cname, "cluster_"); write_number(cname + 8, sizeof cname - 8, (int)j, ngd);
            std::vector<double> locs((max_i - off) * 3 + 1);
            for (size_t q = off; q < max_i; ++q) { locs[3 * (q - off)] = monitor_locs[q].x; locs[3 * (q - off) + 1] = monitor_locs[q].y; locs[3 * (q - off) + 2] = monitor_locs[q].z; }
            h.dataset(std::string(cname) + "/locations", t_loc, locs.data(), max_i - off, 24);
            off = max_i;
            for (; i < max_i; ++i) {
                if (field_times[i].size() < 2) break;
                strcpy(pname, "point_"); write_number(pname + 6, sizeof pname - 6, (int)i, npd);
                const std::string base = std::string(cname) + "/" + pname;
                // sic: t_space holds n_t_pts / save_span samples (disp.cpp:771) although ceil(n_t_pts / save_span) were pushed
                h.dataset(base + "/time", t_cplx, field_times[i].data(), std::min(field_times[i].size(), (size_t)(n_t_pts / save_span)), 16);
                h.dataset(base + "/frequency", t_cplx, &spec[(size_t)i * n_freq * 2], n_freq, 16);
            }
        }
        if (h.save(path)) { printf("cannot write %s\n", path); return -1; }
        printf("finished writing hdf5 file!\n");
    }
    snprintf(path, sizeof path, "%s/field_samples.npz", fname_prefix);
    NpzWriter w(path);
    if (!w.fp) { printf("cannot open %s\n", path); return -1; }
    const size_t n_locs = monitor_locs.size();
    const double ttot_fs = meep_time_to_fs(ttot);
    double tb[3] = {0.0, ttot_fs, n_t_pts ? ttot_fs * save_span / n_t_pts : 0.0};
    w.add("info/time_bounds", tb, 3, 0);
    double v = (double)monitor_clusters.size(); w.add("info/n_clusters", &v, 1, 0);
    v = (double)(n_t_pts / save_span); w.add("info/n_time_points", &v, 1, 0);
    std::vector<double> src(sources.size() * 6);
    for (size_t i = 0; i < sources.size(); ++i) {
        const sj_source_info &q = sources[i];
        const double row[6] = {q.wavelen, q.width, q.phase, q.start_time, q.end_time, q.amplitude};
        memcpy(&src[6 * i], row, sizeof row);
    }
    if (!sources.empty()) w.add("info/sources", src.data(), sources.size(), 6);
    context &c = problem.get_context();
    for (size_t i = c.size(); i > 0; --i) {
        name_val_pair nv = c.peek(i);
        if (nv.get_val().type == VAL_NUM && nv.get_name() && nv.get_name()[0]) {
            double x = nv.get_val().val.x;
            w.add(std::string("info/cgs_params/") + nv.get_name(), &x, 1, 0);
        }
    }
    const size_t ngd = (size_t)(log((double)std::max<size_t>(monitor_clusters.size(), 1)) / log(10.0)) + 1;
    const size_t npd = (size_t)(log((double)std::max<size_t>(n_locs, 1)) / log(10.0)) + 1;
    size_t i = 0, off = 0;
    for (size_t j = 0; j < monitor_clusters.size() + 1; ++j) {          // sic: one empty trailing cluster (disp.cpp:879)
        const size_t max_i = j >= monitor_clusters.size() ? n_locs : monitor_clusters[j];
        char cname[64], pname[64];
        strcpy(cname, "cluster_"); write_number(cname + 8, sizeof cname - 8, (int)j, ngd);
        std::vector<double> locs((max_i - off) * 3);
        for (size_t q = off; q < max_i; ++q) { locs[3 * (q - off)] = monitor_locs[q].x; locs[3 * (q - off) + 1] = monitor_locs[q].y; locs[3 * (q - off) + 2] = monitor_locs[q].z; }
        w.add(std::string(cname) + "/locations", locs.data(), max_i - off, 3);
        off = max_i;
        for (; i < max_i; ++i) {
            if (field_times[i].size() < 2) { printf("Error: monitor location %zu has insufficient points\n", i); break; }
            strcpy(pname, "point_"); write_number(pname + 6, sizeof pname - 6, (int)i, npd);
            w.add(std::string(cname) + "/" + pname + "/time", (const double *)field_times[i].data(), std::min(field_times[i].size(), (size_t)(n_t_pts / save_span)), 2);
        }
    }
    w.close();
    printf("finished writing npz file!\n");
    return 0;
}
