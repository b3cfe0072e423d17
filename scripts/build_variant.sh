#!/bin/bash
# Build a variant of the library with extra -D flags on the TMA kernel units (for A/B measurements on the GPU box):
#   scripts/build_variant.sh NAME "-DSJ_TMA_PROF ..."   ->  sim_juncs_b200/lib/libsimjuncs_b200_NAME.so  (select with SJ_LIB=NAME)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME=$1; DEFS=$2
C=$ROOT/sim_juncs_b200/csrc; B=$ROOT/sim_juncs_b200/_build; O=/tmp/sjvar_$NAME
mkdir -p $O
make -C $C -s -j4 >/dev/null
FLAGS="-O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2,-Wall -Xptxas -v"
# F64_ONLY=1: only the fp64 kernels carry the flags (the fp32 unit and the host unit are the regular build's objects)
UNITS="sj_tma_f64 sj_tma_f32 sj_tma_host"
if [ -n "$F64_ONLY" ]; then UNITS="sj_tma_f64"; cp $B/sj_tma_f32.o $B/sj_tma_host.o $O/; fi
for f in $UNITS; do
  (cd $C && nvcc $FLAGS $DEFS -c $f.cu -o $O/$f.o 2> $O/$f.ptxas.log || (cat $O/$f.ptxas.log; false)) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $ROOT/sim_juncs_b200/lib/libsimjuncs_b200_$NAME.so \
  $B/sj_engine.o $B/sj_launch_f64.o $B/sj_launch_f32.o $O/sj_tma_host.o $O/sj_tma_f64.o $O/sj_tma_f32.o $B/sj_cep.o $B/sj_raster.o -lcudart
grep -A1 "Compiling entry function.*_tma" $O/sj_tma_f64.ptxas.log | grep -E "registers|spill" | head -4
echo built $NAME
