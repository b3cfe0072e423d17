#!/bin/bash
# ncu evidence of round 2, final code (fused step kernel): launch list of the bench command, one full-set capture of step_tma,
# DRAM traffic per launch (fused and two-launch), the two microbenchmarks, and the producer / consumer cycle counters if the
# profiling variant is present (scripts/build_variant.sh prof "-DSJ_TMA_PROF").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"tma|sample_monitors" -s 30 -c 12 --csv \
    --log-file gpurun_out/ncu_launches_r2b.csv python bench.py --steps 20 --warmup 5 --repeats 1 --no-cpu --no-fp32 > gpurun_out/ncu_launches_bench_r2b.log 2>&1
SJ_NO_GRAPH=1 ncu --set full --import-source on --clock-control none --cache-control none -k regex:tma --launch-skip 12 --launch-count 1 -f -o /tmp/ncu_full_r2b \
    python scripts/prof_steps.py 30 > gpurun_out/ncu_full_r2b.log 2>&1
ncu -i /tmp/ncu_full_r2b.ncu-rep --page raw --csv > gpurun_out/ncu_full_r2b_raw.csv 2>> gpurun_out/ncu_full_r2b.log
python scripts/ncu_summary.py gpurun_out/ncu_full_r2b_raw.csv gpurun_out/ncu_full_r2b_summary.json
scripts/ncu_traffic.sh fused_r2b > gpurun_out/traffic_fused_r2b.txt 2>&1
scripts/ncu_traffic.sh twolaunch_r2b SJ_TMA_WAVE=0 > gpurun_out/traffic_twolaunch_r2b.txt 2>&1
timeout 100 scripts/microbench/tma_rate > gpurun_out/microbench_tma_rate.txt 2>&1
timeout 60 scripts/microbench/l2_window > gpurun_out/microbench_l2_window.txt 2>&1
if [ -f sim_juncs_b200/lib/libsimjuncs_b200_prof.so ]; then
  for prec in f64 f32; do
    SJ_LIB=prof SJ_NO_GRAPH=1 python scripts/quick_time.py 300 $prec 2>&1 | tail -2
    SJ_LIB=prof SJ_NO_GRAPH=1 SJ_TMA_WAVE=0 python scripts/quick_time.py 300 $prec 2>&1 | tail -3
  done > gpurun_out/tma_prof_r2b.txt
fi
cat gpurun_out/traffic_fused_r2b.txt gpurun_out/traffic_twolaunch_r2b.txt | tail -12
