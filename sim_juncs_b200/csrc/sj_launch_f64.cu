// f64 instantiation of the half-pass launch logic and of every stepping kernel (see sj_launch.cuh)
#include "sj_launch.cuh"

int sj_launch_pass_f64(sj_sim *s, int which, int k0, int k1, cudaStream_t st) { return launch_pass<double, 2>(s, which, k0, k1, st); }
int sj_profile_f64(sj_sim *s, int reps, double out[4]) { return profile_impl<double, 2>(s, reps, out); }
