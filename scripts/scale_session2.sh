#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out/scale_r2b.jsonl
: > $O
for N in 8 4; do
  timeout 300 $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 20 2> gpurun_out/bench_n${N}_r2b.err | grep '^{' > gpurun_out/bench_n${N}_r2b.json
done
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2952$N scripts/scale_c4.py --res 56 --steps 20 2>> gpurun_out/scale_r2b.err | grep '^{' >> $O
done
timeout 500 $TR --nproc-per-node 8 --master-port 29532 scripts/scale_c4.py --res 88 --steps 10 2>> gpurun_out/scale_r2b.err | grep '^{' >> $O
python - <<'PY'
import json
for N in (4, 8):
    try:
        d = json.load(open("gpurun_out/bench_n%d_r2b.json" % N))
        print("bench N=%d: %.4f ms/step  %.2f G  e2e %.2f G  %s" % (N, d["ms_per_step"], d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["config"]["decomposition"][:60]))
    except Exception as e:
        print("bench N=%d failed: %s" % (N, e))
for l in open("gpurun_out/scale_r2b.jsonl"):
    d = json.loads(l)
    print(d["workload"][:70], "| N", d.get("n_gpus"), "| %.3f ms/step | %.1f G/s" % (d["ms_per_step"], d["cell_updates_per_s"] / 1e9),
          "| roofline %.3f" % d["step_roofline_frac_of_N_x_peak"], "| mem %.0f GB" % d["device_mem_used_GB_max"], "| planes r0", d["planes_rank0"])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/scale_r2b.err | tail -5
