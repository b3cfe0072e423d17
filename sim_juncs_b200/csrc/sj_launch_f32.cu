// f32 instantiation of the half-pass launch logic and of every stepping kernel (see sj_launch.cuh)
#include "sj_launch.cuh"

int sj_launch_pass_f32(sj_sim *s, int which, int k0, int k1, cudaStream_t st) { return launch_pass<float, 4>(s, which, k0, k1, st); }
int sj_profile_f32(sj_sim *s, int reps, double out[4]) { return profile_impl<float, 4>(s, reps, out); }
