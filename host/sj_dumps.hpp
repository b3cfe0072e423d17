// sj_dumps.hpp -- the whole-grid dumps of bound_geom::run (reference src/disp.cpp:696, 732-737) on the C ABI, shared by the
// two C++ hosts (sj_bound_geom.cpp, meep_compat/meep_compat.cpp):
//   fields.output_hdf5(meep::Dielectric, total_volume)          -> <dir>/eps-000000.00.h5, dataset "eps"
//   fields.output_hdf5(meep::Ex, vol.surroundings(), file)      -> <dir>/<name>.h5, datasets "ex.r", "ex.i"
// Values at the pixel centres, x slowest (meep's output grid and axis order, recalled): the mean of the four Yee points of
// the component around each centre; eps = 3 / sum_c <1 / eps_c>.  Same expression order as sim_juncs_b200/output.py, so
// the three hosts write identical files.
#pragma once
#include <string>
#include <vector>

#include "sim_juncs_b200.h"
#include "sj_hdf5.hpp"

namespace sj_dump {

// E-component Yee array [k][j][i] of (n+1)^3 points -> n^3 centred values [i][j][k]
inline std::vector<double> centred(const std::vector<double> &a, const int n[3], int comp) {
    const size_t sx = n[0] + 1, sy = n[1] + 1;
    std::vector<double> out((size_t)n[0] * n[1] * n[2]);
    for (int i = 0; i < n[0]; ++i)
        for (int j = 0; j < n[1]; ++j)
            for (int k = 0; k < n[2]; ++k) {
                const double *p = &a[((size_t)k * sy + j) * sx + i];
                const size_t dj = sx, dk = sx * sy;
                double v;
                if (comp == 0) v = 0.25 * (((p[0] + p[dk]) + p[dj]) + p[dk + dj]);
                else if (comp == 1) v = 0.25 * (((p[0] + p[dk]) + p[1]) + p[dk + 1]);
                else v = 0.25 * (((p[0] + p[dj]) + p[1]) + p[dj + 1]);
                out[((size_t)i * n[1] + j) * n[2] + k] = v;
            }
    return out;
}

// eps_c[c]: eps_inf at the Yee points of E component c
inline int write_eps(const std::string &dir, const std::vector<double> eps_c[3], const int n[3]) {
    std::vector<double> tr((size_t)n[0] * n[1] * n[2], 0.0);
    for (int q = 0; q < 3; ++q) {
        std::vector<double> inv(eps_c[q].size());
        for (size_t i = 0; i < inv.size(); ++i) inv[i] = 1.0 / eps_c[q][i];
        const std::vector<double> cq = centred(inv, n, q);
        for (size_t i = 0; i < tr.size(); ++i) tr[i] = tr[i] + cq[i];
    }
    for (size_t i = 0; i < tr.size(); ++i) tr[i] = 3.0 / tr[i];
    sj_h5::Writer w;
    w.dataset_f64_nd("eps", tr.data(), {(uint64_t)n[0], (uint64_t)n[1], (uint64_t)n[2]});
    return w.save((dir + "/eps-000000.00.h5").c_str());
}

// eps_inf per Yee point from the engine's material table and ids (the rasterizer's output)
inline int write_eps_from_sim(const std::string &dir, sj_sim *sim, const int n[3]) {
    int32_t nm = 0;
    std::vector<sj_material> tab(256);
    if (sj_get_material_table(sim, &nm, tab.data(), 256)) return -1;
    const size_t npt = (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
    std::vector<double> eps_c[3];
    std::vector<uint8_t> ids(npt);
    for (int c = 0; c < 3; ++c) {
        if (sj_get_material_ids(sim, c, ids.data())) return -1;
        eps_c[c].resize(npt);
        for (size_t i = 0; i < npt; ++i) eps_c[c][i] = tab[ids[i]].eps_inf;
    }
    return write_eps(dir, eps_c, n);
}

// one E component of every field set (set 0 -> ".r", set 1 -> ".i")
inline int write_field(const std::string &path, sj_sim *sim, int comp, int n_sets, const int n[3]) {
    static const char *nm[3] = {"ex", "ey", "ez"};
    std::vector<double> raw((size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1));
    sj_h5::Writer w;
    for (int set = 0; set < n_sets && set < 2; ++set) {
        if (sj_get_field(sim, comp, set, raw.data())) return -1;
        const std::vector<double> cq = centred(raw, n, comp);
        w.dataset_f64_nd(std::string(nm[comp]) + (set ? ".i" : ".r"), cq.data(), {(uint64_t)n[0], (uint64_t)n[1], (uint64_t)n[2]});
    }
    return w.save(path.c_str());
}

}  // namespace sj_dump
