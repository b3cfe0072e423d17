#!/bin/bash
# ncu evidence of round 2: launch list of the bench command + one full-set capture of the two step kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"tma|sample_monitors" -s 60 -c 18 --csv \
    --log-file gpurun_out/ncu_launches_r2.csv python bench.py --steps 20 --warmup 5 --repeats 1 --no-cpu > gpurun_out/ncu_launches_bench_r2.log 2>&1
SJ_NO_GRAPH=1 ncu --set full --import-source on --clock-control none -k regex:tma --launch-skip 8 --launch-count 2 -f -o /tmp/ncu_full_r2 \
    python scripts/prof_steps.py 8 > gpurun_out/ncu_full_r2.log 2>&1
ncu -i /tmp/ncu_full_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_r2_raw.csv 2>> gpurun_out/ncu_full_r2.log
python scripts/ncu_summary.py gpurun_out/ncu_full_r2_raw.csv gpurun_out/ncu_full_r2_summary.json
ncu -i /tmp/ncu_full_r2.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -c 30000000 > /tmp/src.csv
python - <<'PY'
import csv, json, collections
# per-kernel DRAM bytes per launch -> roofline.traffic of bench.py
d = json.load(open("gpurun_out/ncu_full_r2_summary.json"))
out = {}
for k, v in d.items():
    name = "h_tma" if k.startswith("h_tma") else "e_tma" if k.startswith("e_tma") else None
    if name:
        out[name] = round((v["dram_read_MB"] + v["dram_write_MB"]) * 1e6)
out["source"] = "ncu --set full, one launch each, SJ_NO_GRAPH=1 python scripts/prof_steps.py 8 (scripts/final_ncu_r2.sh)"
json.dump(out, open("gpurun_out/traffic_r2.json", "w"), indent=1)
print(out)
PY
tail -3 gpurun_out/ncu_launches_bench_r2.log
