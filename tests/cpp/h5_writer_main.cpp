// Writes a small field_samples-shaped file with host/sj_hdf5.hpp; tests/test_hdf5.py compares it byte for byte with
// the file sim_juncs_b200/hdf5.py writes for the same content.
#include <cmath>
#include <vector>
#include "sj_hdf5.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const int n_points = atoi(argv[2]);
    sj_h5::Writer w;
    std::vector<std::pair<std::string, unsigned> > cplx, loc, src;
    cplx.push_back(std::make_pair("Re", 0u)); cplx.push_back(std::make_pair("Im", 8u));
    loc.push_back(std::make_pair("x", 0u)); loc.push_back(std::make_pair("y", 8u)); loc.push_back(std::make_pair("z", 16u));
    const char *sn[6] = {"wavelen", "width", "phase", "start_time", "end_time", "amplitude"};
    for (int i = 0; i < 6; ++i) src.push_back(std::make_pair(std::string(sn[i]), 8u + 8u * i));
    const double tb[3] = {0.0, 300.0, 1.5};
    w.dataset_f64("info/time_bounds", tb, 3);
    const uint64_t nc = 1; w.dataset_u64("info/n_clusters", &nc, 1);
    double s[7] = {0, 0.76, 1.27, 0.25, 5.0, 20.2, 1.0};
    w.dataset("info/sources", sj_h5::compound_type(src, 56), s, 1, 56);
    w.group("info/cgs_params");
    std::vector<double> locs(3 * n_points);
    for (int i = 0; i < 3 * n_points; ++i) locs[i] = 0.25 * i;
    w.dataset("cluster_0/locations", sj_h5::compound_type(loc, 24), locs.data(), n_points, 24);
    w.dataset("cluster_1/locations", sj_h5::compound_type(loc, 24), locs.data(), 0, 24);
    for (int i = 0; i < n_points; ++i) {
        char name[64]; snprintf(name, sizeof name, "cluster_0/point_%04d/time", i);
        std::vector<double> t(2 * 9);
        for (int k = 0; k < 9; ++k) { t[2 * k] = i + 0.5 * k; t[2 * k + 1] = -k; }
        w.dataset(name, sj_h5::compound_type(cplx, 16), t.data(), 9, 16);
    }
    return w.save(argv[1]);
}
