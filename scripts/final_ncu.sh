#!/bin/bash
# ncu evidence only (launch list of the bench command + one full-set step), plus a z-chunk sweep: bash scripts/final_ncu.sh r1c
cd "$(dirname "$0")/.."
R=${1:-r1c}
mkdir -p gpurun_out
run() { env "$@" timeout 100 python scripts/quick_time.py 300 f64 2>&1 | tail -1; }
{ run A=1; run SJ_ZCHUNK=12; run SJ_ZCHUNK=24; run SJ_ZCHUNK=32; run SJ_ZC_FACE=12; run SJ_ZC_FACE=16; run SJ_ZC_FACE=16 SJ_ZCHUNK=24; run SJ_ZC_GEN=8; } > gpurun_out/sweep_$R.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 52 --csv \
    --log-file gpurun_out/ncu_launches_$R.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu_launches_bench_$R.log 2>&1
SJ_NO_GRAPH=1 ncu --set full --import-source on --clock-control none --launch-skip 440 --launch-count 26 -f -o /tmp/ncu_full_$R \
    python scripts/prof_steps.py 24 > gpurun_out/ncu_full_$R.log 2>&1
ncu -i /tmp/ncu_full_$R.ncu-rep --page raw --csv > gpurun_out/ncu_full_${R}_raw.csv 2>> gpurun_out/ncu_full_$R.log
cat gpurun_out/sweep_$R.txt
