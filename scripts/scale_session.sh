#!/bin/bash
# one 8-GPU session: strong scaling of the bench workload, C4 (Au_SiO2_box at --grid-res 56 / 72 / 88), C5 phase sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out/scale_r2.jsonl
: > $O
for N in 8 4 2; do
  timeout 300 $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 20 2> gpurun_out/bench_n${N}_r2.err | grep '^{' > gpurun_out/bench_n${N}_r2.json
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/bench_n1_r2.json
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2952$N scripts/scale_c4.py --res 56 --steps 20 2>> gpurun_out/scale_r2.err | grep '^{' >> $O
done
timeout 400 $TR --nproc-per-node 8 --master-port 29531 scripts/scale_c4.py --res 72 --steps 10 2>> gpurun_out/scale_r2.err | grep '^{' >> $O
timeout 500 $TR --nproc-per-node 8 --master-port 29532 scripts/scale_c4.py --res 88 --steps 10 2>> gpurun_out/scale_r2.err | grep '^{' >> $O
timeout 300 python scripts/phase_sweep.py --n-phases 2 --steps 100 2>> gpurun_out/scale_r2.err | grep '^{' >> $O
timeout 300 $TR --nproc-per-node 8 --master-port 29533 scripts/phase_sweep.py --n-phases 16 --steps 100 2>> gpurun_out/scale_r2.err | grep '^{' >> $O
python - <<'PY'
import json
for N in (1, 2, 4, 8):
    try:
        d = json.load(open("gpurun_out/bench_n%d_r2.json" % N))
        print("bench N=%d: %.4f ms/step  %.2f G  e2e %.2f G" % (N, d["ms_per_step"], d["value"] / 1e9, d["e2e"]["value"] / 1e9))
    except Exception as e:
        print("bench N=%d failed: %s" % (N, e))
for l in open("gpurun_out/scale_r2.jsonl"):
    d = json.loads(l)
    print(d["workload"][:70], "| N", d.get("n_gpus"), "| %.3f ms/step | %.1f G/s" % (d["ms_per_step"], d["cell_updates_per_s"] / 1e9),
          "| roofline %.3f" % d["step_roofline_frac_of_N_x_peak"] if "step_roofline_frac_of_N_x_peak" in d else "", "| mem %.0f GB" % d["device_mem_used_GB_max"] if "device_mem_used_GB_max" in d else "")
PY
tail -5 gpurun_out/scale_r2.err
