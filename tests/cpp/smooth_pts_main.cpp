// Test infrastructure: draws the smoothing offsets with the C++ standard library itself, in the call order of the
// reference's generate_smooth_pts (src/disp.cpp:56-112), and prints them as hex floats.  tests/test_oracle_csg.py
// compares oracle/csg_oracle.c's restatement of seed_seq / mt19937 / generate_canonical / normal_distribution with it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4;
    const double rad = argc > 2 ? atof(argv[2]) : 0.05;
    const uint64_t seed = (uint64_t)(double)0xd9a28bf3;          // DEF_SEED is a double constant passed as uint64_t
    std::seed_seq seeder{(uint32_t)(seed & 0x0000ffff), (uint32_t)((seed & 0xffff0000) >> 32)};
    std::mt19937 gen(seeder);
    std::normal_distribution<double> gaussian(0, rad);
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    for (int i = 0; i < n; ++i) {
        double theta_inv = unif(gen);
        double cos_theta = 1 - 2 * theta_inv;
        double sin_theta = 2 * sqrt(theta_inv * (1 - theta_inv));
        double phi = 2 * M_PI * unif(gen);
        double r = gaussian(gen);
        printf("%a %a %a\n", r * sin_theta * cos(phi), r * sin_theta * cos(phi), r * cos_theta);
    }
    return 0;
}
