#!/bin/bash
# Round-end evidence, one gpurun call on one B200: bench lines (both arms, both precisions), the ncu launch list of the
# bench command itself and one full-set capture of a time step.  Outputs under gpurun_out/ (copied to profiles/ by hand).
cd "$(dirname "$0")/.."
R=${1:-r1b}
mkdir -p gpurun_out
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_f64_$R.json 2> gpurun_out/bench_f64_$R.err
python bench.py --steps 200 --warmup 20 --precision f32 --no-cpu > gpurun_out/bench_f32_$R.json 2> gpurun_out/bench_f32_$R.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 440 -c 44 --csv \
    --log-file gpurun_out/ncu_launches_$R.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu_launches_bench_$R.log 2>&1
SJ_NO_GRAPH=1 ncu --set full --import-source on --clock-control none --launch-skip 380 --launch-count 22 -f -o /tmp/ncu_full_$R \
    python scripts/prof_steps.py 24 > gpurun_out/ncu_full_$R.log 2>&1
ncu -i /tmp/ncu_full_$R.ncu-rep --page raw --csv > gpurun_out/ncu_full_${R}_raw.csv 2>> gpurun_out/ncu_full_$R.log
ls -la /tmp/ncu_full_$R.ncu-rep >> gpurun_out/ncu_full_$R.log
python scripts/trace_step.py > gpurun_out/trace_step_$R.txt 2>&1
tail -c 600 gpurun_out/bench_f64_$R.json; echo; cut -c1-200 gpurun_out/bench_f32_$R.json; echo; cut -c1-200 gpurun_out/bench_ref_$R.json
