// f64 instantiation of the TMA-staged column kernels (see sj_tma.cuh, sj_tma_launch.cuh)
#include "sj_tma_launch.cuh"

int sj_tma_pass_f64(sj_sim *s, int which, cudaStream_t st, bool tick) { return tma_pass<double>(s, which, st, tick); }
