"""BASELINE configs[4]: the quartz_box CEP phase sweep (reference: scripts/run.sh:47-66 job array + scripts/phases.py) as
replicas across GPUs -- every rank runs its share of the pulse phases as ONE batch of field sets that share the rasterized
materials (BoundGeom(phases=...)); no communication between ranks.
  torchrun --nproc-per-node N scripts/phase_sweep.py --n-phases 16 [--steps K | --full --out-dir DIR]
--steps K times K steps (CUDA events, max over ranks) and prints one JSON line; --full runs every phase to the end and
writes <out-dir>/phase_<index>/field_samples.h5 exactly as a single-phase run would."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-phases", type=int, default=16)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--out-dir", default="/tmp/phase_sweep")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--scene", default="quartz_box")
    args = ap.parse_args()
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.settings import settings_from
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    d = os.path.join(ROOT, "scenes", args.scene)
    st = settings_from(os.path.join(d, "params.conf"))
    st.geom_fname = os.path.join(d, os.path.basename(st.geom_fname or "junc.geom"))
    phases = [2 * math.pi * i / args.n_phases for i in range(args.n_phases)]
    mine = list(range(rank, args.n_phases, world))
    n = st.grid_cells()
    # every phase with its quadrature (phi + pi/2 = the imaginary part of meep's complex field for phi): 2 sets per phase,
    # the work the reference does per phase
    bg = BoundGeom(st, None, precision=args.precision, device=local,
                   phases=[phases[i] for i in mine] + [phases[i] + math.pi / 2 for i in mine])

    def mx(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.full:
        t0 = time.time()
        bg.run()
        t_run = mx(time.time() - t0)
        from sim_juncs_b200.output import save_phase_batch
        save_phase_batch(bg, args.out_dir, mine)
        if rank == 0:
            print(json.dumps({"workload": "%s %d^3 cells, %d phases over %d GPUs, full run of %d steps" % (args.scene, n, args.n_phases, world, bg.n_t_pts),
                              "run_s": t_run, "cell_updates_per_s": float(n) ** 3 * 2 * args.n_phases * bg.n_t_pts / t_run}))
    else:
        bg.sim.run(20, st.save_span or 1)
        torch.cuda.synchronize()
        times = []
        for rep in range(3):
            if world > 1:
                dist.barrier()
            times.append(mx(bg.sim.run_timed(args.steps, st.save_span or 1)))
        ms = sorted(times)[1]
        if rank == 0:
            print(json.dumps({"workload": "%s %d^3 cells, %d CEP phases x 2 quadratures as field sets, %d phases per GPU" % (args.scene, n, args.n_phases, len(mine)),
                              "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
                              "cell_updates_per_s": float(n) ** 3 * 2 * args.n_phases * args.steps / (ms * 1e-3),
                              "phase_steps_per_s": args.n_phases * args.steps / (ms * 1e-3)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
