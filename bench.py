#!/usr/bin/env python
"""bench.py -- Yee cell-updates/s of the FDTD hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f64|f32]

A "step" is one full Yee time step (H-pass + E-pass incl. PML, ADE and source) of every field set over the whole
grid, monitors sampled every save_span steps as the reference does (src/disp.cpp:719-741).

Workload at every N: BASELINE configs[1], junctions/Au_graphene_box at production resolution -- 181^3 cells, complex
fields = 2 field sets, Drude-Lorentz Au + graphene sheet + Lorentz SiO2, fp64 -- entered the way a user enters it:
scenes/Au_graphene_box/params.conf + junc.geom through the package's own conf / CGS readers and BoundGeom.
  N = 1 : the whole box on one GPU.
  N > 1 : the SAME box cut into N z-slabs, one rank per GPU (strong scaling).  The slabs are connected inside the
          library (CUDA-IPC peer memory): the step kernels write a slab's boundary plane straight into the neighbour's
          halo over NVLink and the slabs order themselves through device-side flags; no collective per step.
`value`: K steps timed with CUDA events on the simulation's stream, bracketed by barrier + synchronize, max over ranks;
the region is repeated (--repeats, default 5) and the median is reported.  `e2e`: the same K steps through the public
API by the host clock, with the source drive table uploaded from the host and every monitor sample read back to the
host inside the timed region (max over ranks).  `--impl reference` times the CPU oracle (meep-structured C/OpenMP
restatement; meep itself is not installable here) on all host cores on the same workload; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"          # the version banner goes to stdout, where the one JSON line belongs

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SCENE = "Au_graphene_box"
SAVE_SPAN = 20
# the same `config` object in both arms (the driver compares them)
WORKLOAD = "junctions/Au_graphene_box 181^3 cells x 2 field sets (complex), res 10 -> 181/18"
CONFIG = {"workload": WORKLOAD, "scene": "scenes/Au_graphene_box/params.conf + junc.geom (re-authored, see DESIGN.md) through the own conf/CGS readers",
          "save_span": SAVE_SPAN, "l2": "fields + auxiliaries of a step (2.4 GB) >> 126 MB L2; no flush needed"}


def load_settings():
    """params.conf of the scene through the package's own reader (argparse.h semantics), as `sim_geom` would"""
    from sim_juncs_b200.settings import settings_from
    d = os.path.join(ROOT, "scenes", SCENE)
    s = settings_from(os.path.join(d, "params.conf"))
    s.geom_fname = os.path.join(d, os.path.basename(s.geom_fname or "junc.geom"))
    s.save_span = SAVE_SPAN
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
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def oracle_problem(st):
    """The same workload for the CPU oracle (masks from the C restatement of the CSG tests)."""
    from helpers import oracle_bound_geom, oracle_raster
    from sim_juncs_b200.scene import Scene
    sc = Scene.from_geom(st.geom_fname, st)
    masks = [oracle_raster(sc, st, c) for c in range(3)]
    o, n_t = oracle_bound_geom(sc, st, masks, nsets=2)
    return o


def host_threads():
    """all host cores, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)          # read by libgomp when the oracle library is loaded
    return n


def run_reference(args):
    """CPU arm: the oracle port on all host threads, same workload / metric / unit as the repo arm at this N (the
    workload does not depend on N: strong scaling).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    want = host_threads()
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(CONFIG),
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port", "threads_requested": want,
                             "sample": "%d full-grid oracle steps (meep-structured C/OpenMP restatement; meep itself is absent)" % args.steps},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def plane_weights(n, pml_cells):
    """relative bytes per z-plane: planes inside the z-PML carry the UPML auxiliaries of every cell"""
    return [1.22 if (k < pml_cells or k > n - pml_cells) else 1.0 for k in range(n + 1)]


def run_ours(args):
    import torch
    from sim_juncs_b200.bound_geom import BoundGeom
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # the version banner (env or nccl.conf) goes to stdout, where the one JSON line belongs
        import torch.distributed as dist
        # NCCL writes its version banner to stdout when the communicator is created; stdout carries the one JSON line, so
        # file descriptor 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    st = load_settings()
    n = st.grid_cells()
    prec = args.precision
    esz = 8 if prec == "f64" else 4
    n_sets = 2
    peak, peak_src = measured_peak()
    sampler = ClockSampler(local)
    K, W = args.steps, args.warmup
    cells = float(n) ** 3

    kz = None
    if world > 1:
        from sim_juncs_b200.parallel import balanced_slab, connect_slabs
        # slabs of equal bytes per step, from the rasterized scene (substrate planes carry polarisation traffic)
        kz = balanced_slab(lambda kz_: BoundGeom(st, None, precision=prec, n_sets=n_sets, device=local, kz=kz_), n + 1, rank, world)
        all_kz = [None] * world
        dist.all_gather_object(all_kz, kz)

    def make():
        bg = BoundGeom(st, None, precision=prec, n_sets=n_sets, device=local, kz=kz)
        if world > 1:
            connect_slabs(bg.sim, rank, world)
        return bg

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-timed: K steps, CUDA events on the simulation's stream, repeated; median of the max over ranks ----
    bg = make()
    sim = bg.sim
    sim.run(W, SAVE_SPAN)
    barrier()
    if rank == 0:
        sampler.start()
    times = []
    l1 = sim.launches()
    for rep in range(args.repeats):
        barrier()
        ms = sim.run_timed(K, SAVE_SPAN)            # synchronises the stream on both sides of the K steps
        barrier()
        times.append(max_over_ranks(ms))
    launches = (sim.launches() - l1) // args.repeats
    extended = False
    if len(sampler.samples) < 6 or world > 1:       # keep the same load running while nvidia-smi gets its samples
        extended = True
        t_end = time.time() + 2.5
        flag, extra = 1, 0
        while flag and extra < 2400:      # bounded in steps too: this scene's fields start to grow after ~6000 steps
            sim.run(60, SAVE_SPAN)
            extra += 60
            flag = int(max_over_ranks(1.0 if time.time() < t_end else 0.0)) if world > 1 else int(time.time() < t_end)
    sampler.stop_flag = True
    ms = statistics.median(times)
    value = cells * n_sets * K / (ms * 1e-3)
    bytes_step = max_over_ranks(sim.bytes_per_step()) if world == 1 else None
    del bg, sim

    # ---- end to end through the public API with host buffers: fresh simulation, K steps by the host clock ----
    bg2 = make()
    bg2.sim.run(W, SAVE_SPAN)
    up0 = bg2.sim.h2d_bytes()
    barrier()
    t0 = time.perf_counter()
    done = 0
    while done < K:                                   # BoundGeom.run's own chunking (save_span steps per call)
        chunk = min(SAVE_SPAN, K - done)
        bg2.sim.run(chunk, SAVE_SPAN, sync=False)
        done += chunk
    mon = bg2.sim.monitors()                          # D2H of every sample of this rank (synchronises)
    bg2.sim.sync()
    te = max_over_ranks(time.perf_counter() - t0)
    h2d = bg2.sim.h2d_bytes() - up0
    e2e = {"value": cells * n_sets * K / te, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / K,
           "d2h_bytes_per_step": mon.nbytes / K,
           "note": "BoundGeom/sj_run in save_span chunks by the host clock (max over ranks); inside the timed region: the host-"
                   "evaluated source drive table for these steps is uploaded (%d B, the warm-up table does not reach that far) "
                   "and the monitor series is read back to the host" % int(h2d)}
    del bg2

    roofline, cpu, per_rank, fp32 = None, None, None, None
    if world == 1:
        # ---- per-kernel roofline, timed live with CUDA events on a scratch simulation ----
        scratch = make().sim
        scratch.run(40, SAVE_SPAN)
        prof = scratch.profile_kernels(reps=20)
        cnt = scratch.counts()
        alg = {"h_pass": n_sets * (cnt["cells"] * 9 * esz + 4 * esz * cnt["pml_cells"]),
               "e_pass": n_sets * (cnt["cells"] * (9 * esz + 3) + 3 * esz * cnt["pole_points"] + 4 * esz * cnt["pml_cells"])}
        t_ms = {"h_pass": prof["h_interior"], "e_pass": prof["e_interior"]}
        fused_ms = prof["h_pml"]            # > 0: the step is ONE launch (fused H + E wavefront kernel); out[2] of sj_profile_kernels
        kern = {k: {"kernel": "step_tma(H items)" if k == "h_pass" else "step_tma(E items)", "ms": t_ms[k], "algorithmic_bytes": alg[k],
                    "achieved_gbs": alg[k] / (t_ms[k] * 1e-3) / 1e9, "frac": alg[k] / (t_ms[k] * 1e-3) / 1e9 / peak} for k in alg}
        if fused_ms > 0:
            b = bytes_step                  # sj_bytes_per_step: both half-passes, device-side counts
            kern["step"] = {"kernel": "step_tma", "ms": fused_ms, "algorithmic_bytes": b, "achieved_gbs": b / (fused_ms * 1e-3) / 1e9,
                            "frac": b / (fused_ms * 1e-3) / 1e9 / peak}
            dom = "step"
        else:
            dom = max(kern, key=lambda k: kern[k]["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_r2b.json")
        if prec == "f64" and os.path.exists(tp):
            traffic = json.load(open(tp)).get(kern[dom]["kernel"])
        step_gbs = bytes_step / (ms * 1e-3 / K) / 1e9
        roofline = {"bound": "hbm", "kernel": kern[dom]["kernel"], "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes"], "kernel_ms": kern[dom]["ms"],
                    "h_pass_ms": t_ms["h_pass"], "h_pass_bytes": alg["h_pass"], "h_pass_frac": kern["h_pass"]["frac"],
                    "e_pass_ms": t_ms["e_pass"], "e_pass_bytes": alg["e_pass"], "e_pass_frac": kern["e_pass"]["frac"],
                    "step_algorithmic_bytes": bytes_step, "step_achieved_gbs": step_gbs, "step_frac": step_gbs / peak,
                    "families": "one persistent TMA kernel (step_tma) covers interior, PML-face and PML-edge cells of both half-passes; "
                                + ("a step is ONE launch of it (fused z wavefront: the E-pass items find the H planes just written in the L2, so "
                                   "DRAM traffic is below the algorithmic bytes); h_pass / e_pass are the same kernel given only the items of "
                                   "one half-pass, timed alone" if fused_ms > 0 else
                                   "a step is two launches of it (H-pass items, then E-pass items)")}
        del scratch
        # ---- the same workload in fp32 (the reference tolerates 1e-4: north_star's second precision), device-timed, short ----
        fp32 = None
        if prec == "f64" and not args.no_fp32:
            b32 = BoundGeom(st, None, precision="f32", n_sets=n_sets, device=local)
            b32.sim.run(W, SAVE_SPAN)
            ms32 = statistics.median(b32.sim.run_timed(100, SAVE_SPAN) for _ in range(3)) / 100
            by32 = b32.sim.bytes_per_step()
            fp32 = {"ms_per_step": ms32, "value": cells * n_sets / (ms32 * 1e-3), "unit": "cell-updates/s",
                    "step_algorithmic_bytes": by32, "step_frac": by32 / (ms32 * 1e-3) / 1e9 / peak}
            del b32
        # ---- CPU baseline: the oracle on a bounded sample of the same workload ----
        if not args.no_cpu:
            host_threads()
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
    if rank == 0:
        cfg = dict(CONFIG)
        decomposition = None
        if world > 1:
            decomposition = "%d z-slabs of the same box, %s planes (equal bytes per step); boundary planes written into the neighbour's halo by the step " \
                                   "kernels (peer stores over NVLink, CUDA-IPC mapped), device-side flags, one CUDA graph per step and rank" \
                                   % (world, "/".join(str(b - a) for a, b in all_kz))
        line = {"metric": "yee_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": prec, "data": "synthetic", "config": cfg,
                "clocks": dict(sampler.summary(), extended_sampling=extended), "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu,
                "repeats": args.repeats, "ms_per_step_all_repeats": [t / K for t in times], "decomposition": decomposition,
                "fp32": fp32}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fp32", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.repeats = max(args.repeats, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
