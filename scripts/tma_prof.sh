#!/bin/bash
# ncu capture of the TMA kernels of one step + knob sweep: bash scripts/tma_prof.sh <tag>
cd "$(dirname "$0")/.."
R=${1:-t1}
mkdir -p gpurun_out
SJ_NO_GRAPH=1 SJ_TMA=3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:tma --launch-skip 8 --launch-count 2 -f -o /tmp/ncu_tma_$R \
    python scripts/prof_steps.py 8 > gpurun_out/ncu_tma_$R.log 2>&1
ncu -i /tmp/ncu_tma_$R.ncu-rep --page raw --csv > gpurun_out/ncu_tma_${R}_raw.csv 2>> gpurun_out/ncu_tma_$R.log
python scripts/ncu_summary.py gpurun_out/ncu_tma_${R}_raw.csv gpurun_out/ncu_tma_${R}_summary.json
run() { env "$@" timeout 100 python scripts/tma_check.py time f64 env 2>&1 | grep "^time"; }
{ run SJ_TMA=0; run A=1; run A=2; run SJ_TMA_RING_KB=112; run SJ_TMA_RING_KB=144; run SJ_TMA_RING_KB=176; run SJ_TMA_RING_KB=208; run SJ_TMA_ZC=12; run SJ_TMA_ZC=20; run SJ_TMA_FINE=0; run SJ_TMA_FINE=3; run SJ_TMA_WGEN=1; run SJ_TMA=0; } > gpurun_out/tma_sweep_$R.txt 2>&1
cat gpurun_out/tma_sweep_$R.txt
