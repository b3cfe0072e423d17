"""BASELINE configs[3]: junctions/Au_SiO2_box at --grid-res R (56 -> 1009^3, 72 -> 1297^3, 88 -> 1585^3 cells), fp64 x 2 field
sets, z-slabs across the ranks of a torchrun launch (or one GPU).  Prints one JSON line per run (rank 0).
  torchrun --nproc-per-node N scripts/scale_c4.py --res 56 [--steps 20] [--precision f64] [--sets 2]
The scene is the reference parser's output of the shipped junctions/Au_SiO2_box/junc.geom (scenes/json/Au_SiO2_box.json)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=float, default=56.0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--sets", type=int, default=2)
    args = ap.parse_args()
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.settings import ParseSettings
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    st = ParseSettings()
    fixture = os.path.join(ROOT, "scenes", "json", "Au_SiO2_box.json")
    for k, v in json.load(open(fixture))["settings"].items():
        setattr(st, k, v)
    st.grid_num = -1
    st.resolution = args.res
    st.correct_defaults()
    st.save_span = 20
    n = st.grid_cells()
    kz = None
    if world > 1:
        from sim_juncs_b200.parallel import balanced_slab, connect_slabs
        kz = balanced_slab(lambda kz_: BoundGeom(st, fixture, precision=args.precision, n_sets=args.sets, device=local, kz=kz_), n + 1, rank, world)
    t0 = time.time()
    bg = BoundGeom(st, fixture, precision=args.precision, n_sets=args.sets, device=local, kz=kz)
    if world > 1:
        connect_slabs(bg.sim, rank, world)
    t_setup = time.time() - t0
    sim = bg.sim

    def mx(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sm(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sim.run(args.warmup, 20)
    torch.cuda.synchronize()
    times = []
    for rep in range(3):
        if world > 1:
            dist.barrier()
        ms = sim.run_timed(args.steps, 20)
        times.append(mx(ms))
    ms = sorted(times)[1]
    mem = sim.memory()
    free, total = torch.cuda.mem_get_info(local)
    bytes_step = sm(sim.bytes_per_step())
    peak = 6558.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    m = sim.monitors()
    finite = bool(np.isfinite(m).all())
    line = {"workload": "junctions/Au_SiO2_box --grid-res %g: %d^3 = %.3g cells x %d field sets, %s" % (args.res, n, float(n) ** 3, args.sets, args.precision),
            "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "ms_per_step_repeats": [t / args.steps for t in times],
            "cell_updates_per_s": float(n) ** 3 * args.sets * args.steps / (ms * 1e-3),
            "step_algorithmic_bytes": bytes_step, "step_roofline_frac_of_N_x_peak": bytes_step / (ms * 1e-3 / args.steps) / 1e9 / (peak * world),
            "setup_s_max": mx(t_setup), "raster_s_max": mx(bg.t_raster),
            "device_mem_used_GB_max": mx((total - free) / 1e9), "slab_bytes_GB_max": {k: mx(v / 1e9) for k, v in mem.items() if k != "pol_planes"},
            "pol_planes_max": int(mx(mem["pol_planes"])), "planes_rank0": (kz[1] - kz[0]) if kz else n + 1, "monitors_finite": finite}
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
