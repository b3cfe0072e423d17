#!/usr/bin/env python
"""bench.py -- Yee cell-updates/s of the FDTD hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f64|f32]

A "step" is one full Yee time step (H-pass + E-pass incl. PML, ADE and source) of every field set
over the whole grid, monitors sampled every save_span steps as the reference does.
  N = 1 : BASELINE configs[1], junctions/Au_graphene_box at production resolution (181^3 cells,
          complex fields = 2 field sets, Drude Au + graphene sheet + Lorentz SiO2), scene fixture
          scenes/json/Au_graphene_box.json.
  N > 1 : the same scene in a box N times taller (z-slab per rank, weak scaling), one rank per GPU,
          one-plane halo exchange of the tangential fields per half step over NCCL.
`value` is timed with CUDA events with everything resident in HBM; `e2e` runs the same steps through
the public BoundGeom API with host buffers (source table upload + monitor read-back per save inside
the timed region).  `--impl reference` times the CPU oracle (meep-structured restatement; meep itself
is not installable here) on the host cores on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SCENE = "Au_graphene_box"
SAVE_SPAN = 20


def load_settings():
    with open(os.path.join(ROOT, "scenes", "json", SCENE + ".json")) as fp:
        st = json.load(fp)["settings"]
    from sim_juncs_b200.settings import ParseSettings
    s = ParseSettings()
    for k, v in st.items():
        setattr(s, k, v)
    s.out_dir = "/tmp"
    return s


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.samples = []

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def oracle_problem(st, nz_mult=1):
    """The same workload for the CPU oracle (masks from the C restatement of the CSG tests)."""
    from helpers import oracle_bound_geom, oracle_raster
    from sim_juncs_b200.scene import Scene
    sc = Scene.load(os.path.join(ROOT, "scenes", "json", SCENE + ".json"))
    masks = [oracle_raster(sc, st, c) for c in range(3)]
    o, n_t = oracle_bound_geom(sc, st, masks, nsets=2)
    return o


def run_reference(args):
    """CPU arm: the oracle port on all host threads, same config/metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    st = load_settings()
    n = st.grid_cells()
    o = oracle_problem(st)
    cores = orc.lib().orc_num_threads()
    for _ in range(args.warmup):
        o.step()
    t0 = time.time()
    for _ in range(args.steps):
        o.step()
    dt = time.time() - t0
    val = float(n) ** 3 * 2 * args.steps / dt
    line = {"impl": "reference", "metric": "yee_cell_updates_per_s", "value": val, "unit": "cell-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "junctions/Au_graphene_box 181^3 cells x 2 field sets (complex), res 10 -> 181/18",
                       "scene": "scenes/Au_graphene_box/junc.geom (re-authored, see DESIGN.md)"},
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                             "sample": "%d full-grid oracle steps (meep-structured C/OpenMP restatement; meep itself is absent)" % args.steps},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cuda_tensor_from_ptr(torch, ptr, nbytes, device):
    class _P:
        pass
    p = _P()
    p.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(p, device=device)


def run_ours(args):
    import torch
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.engine import Sim
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # the version banner goes to stdout, where the one JSON line belongs
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    st = load_settings()
    n = st.grid_cells()
    prec = args.precision
    esz = 8 if prec == "f64" else 4
    n_sets = 2
    peak, peak_src = measured_peak()
    scene_path = os.path.join(ROOT, "scenes", "json", SCENE + ".json")
    sampler = ClockSampler(local)

    if world == 1:
        bg = BoundGeom(st, scene_path, precision=prec, n_sets=n_sets, device=local)
        sim = bg.sim
        cells = float(n) ** 3
        l0 = sim.launches()
        sim.run(args.warmup, SAVE_SPAN)
        sampler.start()
        l1 = sim.launches()
        ms = sim.run_timed(args.steps, SAVE_SPAN)
        launches = sim.launches() - l1
        extended = False
        if len(sampler.samples) < 4:          # timed region shorter than a few nvidia-smi polls: keep the same load running
            extended = True
            t_end = time.time() + 1.5
            while time.time() < t_end:
                sim.run(50, SAVE_SPAN)
        sampler.stop_flag = True
        value = cells * n_sets * args.steps / (ms * 1e-3)
        bytes_step = sim.bytes_per_step()
        # ---- dominant kernel roofline, timed live with CUDA events on a scratch simulation ----
        scratch = BoundGeom(st, scene_path, precision=prec, n_sets=n_sets, device=local).sim
        scratch.run(40, SAVE_SPAN)
        prof = scratch.profile_kernels(reps=10)
        cnt = scratch.counts()
        alg_e = n_sets * (cnt["interior_cells"] * (9 * esz + 3) + 3 * esz * cnt["pole_points_interior"])
        alg_h = n_sets * cnt["interior_cells"] * 9 * esz
        dom = "e_interior" if prof["e_interior"] >= prof["h_interior"] else "h_interior"
        alg = alg_e if dom == "e_interior" else alg_h
        achieved = alg / (prof[dom] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if prec == "f64" and os.path.exists(tp):
            traffic = json.load(open(tp))["traffic_bytes_per_pass"].get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "traffic": traffic, "kernel_ms": prof,
                    "algorithmic_bytes_per_launch": alg,
                    "step": {"algorithmic_bytes": bytes_step, "achieved_gbs": bytes_step / (ms * 1e-3 / args.steps) / 1e9,
                             "frac": bytes_step / (ms * 1e-3 / args.steps) / 1e9 / peak}}
        del scratch
        # ---- end to end through the public API with host buffers ----
        bg2 = BoundGeom(st, scene_path, precision=prec, n_sets=n_sets, device=local)
        bg2.sim.run(args.warmup, SAVE_SPAN)
        host = torch.empty((args.steps // SAVE_SPAN + 2, bg2.sim.n_mon, n_sets), dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = 0
        d2h = 0
        while done < args.steps:
            chunk = min(SAVE_SPAN, args.steps - done)
            bg2.sim.run(chunk, SAVE_SPAN, sync=False)
            done += chunk
        m = bg2.sim.monitors()                # D2H of every sample taken in the timed region
        d2h = m.nbytes
        bg2.sim.sync()
        te = time.perf_counter() - t0
        e2e_val = cells * n_sets * args.steps / te
        n_src = 1
        e2e = {"value": e2e_val, "unit": "cell-updates/s", "h2d_bytes_per_step": n_src * n_sets * 2 * esz,
               "d2h_bytes_per_step": d2h / max(args.steps, 1),
               "note": "BoundGeom/sj_run in save_span chunks: host-computed source table uploaded (H2D), monitor series read back to host (D2H) inside the timed region"}
        del host
        # ---- CPU baseline: the oracle on a bounded sample of the same workload ----
        cpu = None
        if not args.no_cpu:
            from oracle import oracle as orc
            o = oracle_problem(st)
            o.step()
            t0 = time.time()
            k = 0
            while time.time() - t0 < 12.0 and k < 50:
                o.step()
                k += 1
            tc = time.time() - t0
            cpu = {"value": cells * 2 * k / tc, "unit": "cell-updates/s", "cores": orc.lib().orc_num_threads(), "kind": "port",
                   "sample": "%d full-grid steps of the same 181^3 x 2-set workload (oracle/fdtd_oracle.c, OpenMP)" % k}
        line = {"metric": "yee_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": prec, "data": "synthetic",
                "config": {"workload": "junctions/Au_graphene_box 181^3 cells x 2 field sets (complex), res 10 -> 181/18",
                           "scene": "scenes/Au_graphene_box/junc.geom (re-authored, see DESIGN.md)", "save_span": SAVE_SPAN,
                           "l2": "working set %.0f MB per step >> 126 MB L2 (no flush needed)" % (bytes_step / 1e6)},
                "clocks": dict(sampler.summary(), extended_sampling=extended), "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ N > 1: z-slabs, weak scaling
    st_tall_n2 = n * world                     # box world x taller in z
    from sim_juncs_b200.parallel import slab_range
    kz = slab_range(st_tall_n2 + 1, rank, world)
    from sim_juncs_b200.scene import LIGHT_SPEED, Scene
    sc = Scene.load(scene_path)
    sim = Sim((n, n, st_tall_n2), st.resolution, pml=st.pml_thickness, precision=prec, n_sets=n_sets, kz=kz, device=local)
    nodes = sc.node_array()
    regions = []
    for reg in sc.regions:
        thick = 1.0 / st.resolution if reg.make_2d else 1.0
        if reg.make_2d:
            for i in range(3):
                nodes[reg.root].M[3 * i + 2] = nodes[reg.root].M[3 * i + 2] * thick
        regions.append((reg.root, reg.eps if reg.eps is not None else st.ambient_eps,
                        [(w0 / st.um_scale, g / st.um_scale, sg / thick, not ud) for (w0, g, sg, ud) in reg.poles_raw]))
    sim.rasterize(st.ambient_eps, nodes, regions)
    info, (p1, p2) = sc.sources[0], sc.source_boxes[0]
    cba = LIGHT_SPEED * st.um_scale
    sim.add_gaussian_source(info.component, p1, p2, info.amplitude, 1 / (info.wavelen * st.um_scale), info.width * cba,
                            info.phase, info.start_time * cba, info.end_time * cba, True)
    sim.add_monitors(np.array(sc.monitor_locs), comp=0)
    from sim_juncs_b200.parallel import SlabRunner
    runner = SlabRunner(sim, kz, n_sets, dev, save_span=SAVE_SPAN, overlap=not args.no_overlap)
    stream = runner.tstream                    # the stream the kernels and the NCCL ordering live on

    def step(i):
        runner.step()

    for i in range(args.warmup):
        step(i)
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    l1 = sim.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0 and len(sampler.samples) < 4:
        t_end = time.time() + 1.5
        flag = torch.ones(1, device=dev)
    else:
        t_end = 0.0
        flag = torch.zeros(1, device=dev)
    dist.broadcast(flag, 0)
    if flag.item() > 0:                       # keep the same load running while nvidia-smi samples
        for i in range(300):
            step(args.warmup + args.steps + i)
        torch.cuda.synchronize()
    sampler.stop_flag = True
    ms = float(ms.item())
    cells = float(n) * n * st_tall_n2
    value = cells * n_sets * args.steps / (ms * 1e-3)
    # ---- end to end: the same K steps by the host clock, each rank's monitor series read back to the host
    #      (D2H) inside the timed region; max over ranks
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    runner.synchronize()
    mon = sim.monitors()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = cells * n_sets * args.steps / float(te.item())
    launches = (sim.launches() - l1) * args.steps // max(runner.i - args.warmup, 1)
    if rank == 0:
        line = {"metric": "yee_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": prec, "data": "synthetic",
                "config": {"workload": "Au_graphene_box scene in a box %dx taller in z: %dx%dx%d cells x 2 field sets, one z-slab per GPU" % (world, n, n, st_tall_n2),
                           "halo": "Hx,Hy up / Ex,Ey down, one plane per half step, NCCL send/recv" +
                                   ("" if args.no_overlap else ", boundary plane first and on the wire while the rest of the slab runs"),
                           "save_span": SAVE_SPAN},
                "clocks": sampler.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "cell-updates/s", "h2d_bytes_per_step": 2 * n_sets * 2 * esz,
                        "d2h_bytes_per_step": mon.nbytes / max(runner.i, 1),
                        "note": "SlabRunner.step loop by the host clock (max over ranks), source table uploads and the read-back of every rank's monitor series inside the timed region"},
                "halo_bytes_per_step_per_rank": runner.halo.bytes_per_step(),
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange the halos after each full half-pass")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
