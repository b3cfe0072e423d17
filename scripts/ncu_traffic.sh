#!/bin/bash
# DRAM bytes per launch of the step kernels (ncu, one metric pass, caches left alone): scripts/ncu_traffic.sh TAG [ENV=VAL ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
env "$@" SJ_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:"tma" -s 12 -c 6 --csv --log-file gpurun_out/ncu_traffic_$tag.csv python scripts/prof_steps.py 30 > gpurun_out/ncu_traffic_$tag.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/ncu_traffic_$tag.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:24]), {})[r[mi]] = float(r[vi].replace(",", ""))
for (i, k), m in d.items():
    print("$tag", i, k, "time us %.1f  read MB %.1f  write MB %.1f" % (m.get("gpu__time_duration.sum", 0) / 1e3, m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6))
PY
