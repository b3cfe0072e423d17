"""TMA-staged kernels against the register kernels: bitwise field comparison on a small dispersive + PML box and on the
bench workload, then step timings.  usage: python scripts/tma_check.py [small|bench|time] ...
Each simulation is created with SJ_TMA set in the environment (read by sj_create)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sim_juncs_b200 import Sim  # noqa: E402
from sim_juncs_b200.materials import materials_from_regions  # noqa: E402


def small_sim(mode, prec="f64", n=(40, 36, 44), nsets=2):
    os.environ["SJ_TMA"] = str(mode)
    a = 6.0
    shape = (n[2] + 1, n[1] + 1, n[0] + 1)
    k, j, i = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    masks = [((k > 22).astype(np.uint8) | ((i < 17) & (k > 14) & (k <= 22)).astype(np.uint8) << 1) for _ in range(3)]
    regs = (1.0, [2.25, 1.0], [[(1.1, 0.05, 1.3, 0)], [(1e-10, 0.04, 2.0e19, 1), (0.9, 0.1, 0.7, 0)]])
    g = Sim(n, a, pml=1.0, n_sets=nsets, device=0, precision=prec)
    g.set_materials(materials_from_regions(*regs), masks)
    g.add_gaussian_source(0, [0, 0, 1.5], [n[0] / a, n[1] / a, 1.5], 1.0, 0.4, 1.5, 0.3, 0.0, 18.0, True)
    g.add_monitors([[2.0, 1.6, 2.3], [1.0, 2.0, 3.0]], 0)
    return g


def compare(ga, gb, tag):
    worst = 0.0
    for c in range(6):
        for q in range(ga.n_sets):
            fa, fb = ga.field(c, q), gb.field(c, q)
            d = np.abs(fa - fb).max()
            ref = np.abs(fa).max()
            if d != 0.0:
                bad = np.argwhere(np.abs(fa - fb) > 0)
                print("  %s comp %d set %d: max|diff| %.3e (max|f| %.3e), %d points differ, first (k,j,i) %s" %
                      (tag, c, q, d, ref, len(bad), bad[0]))
            worst = max(worst, d / ref if ref > 0 else d)
    print("%s: worst relative difference %.3e, max|Ex| %.3e" % (tag, worst, np.abs(ga.field(0, 0)).max()))
    return worst


def run_small(prec):
    ref = small_sim(0, prec)
    ref.run(120, 4)
    ok = True
    for mode in (1, 2, 3):
        g = small_sim(mode, prec)
        g.run(120, 4)
        w = compare(ref, g, "small %s SJ_TMA=%d" % (prec, mode))
        ok &= (w == 0.0)
        g.close()
    ref.close()
    return ok


def bench_sim(mode, prec="f64"):
    import bench
    from sim_juncs_b200.bound_geom import BoundGeom
    os.environ["SJ_TMA"] = str(mode)
    st = bench.load_settings()
    return BoundGeom(st, os.path.join(bench.ROOT, "scenes", "json", bench.SCENE + ".json"), precision=prec, n_sets=2)


def run_bench_compare(prec):
    a, b = bench_sim(0, prec), bench_sim(3, prec)
    # move the pulse to t = 0 so the fields are non-zero after a few hundred steps: reuse the run as it is but longer
    for g in (a, b):
        g.sim.run(700, 20)
    w = compare(a.sim, b.sim, "bench %s SJ_TMA=3" % prec)
    return w == 0.0


def run_time(prec, modes=(0, 1, 2, 3), steps=200):
    for mode in modes:
        g = bench_sim(mode, prec)
        g.sim.run(40, 20)
        g.sim.sync()
        best = min(g.sim.run_timed(steps, 20) for _ in range(3)) / steps
        knobs = {k: v for k, v in os.environ.items() if k.startswith("SJ_")}
        print("time %s SJ_TMA=%d: %.4f ms/step  launches/step %.1f  %s" % (prec, mode, best, g.sim.launches() / (40 + 3 * steps), knobs))
        sys.stdout.flush()
        g.sim.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    prec = sys.argv[2] if len(sys.argv) > 2 else "f64"
    t0 = time.time()
    if what == "small":
        ok = run_small(prec)
        print("small %s: %s (%.1f s)" % (prec, "BITWISE EQUAL" if ok else "DIFFERENT", time.time() - t0))
        sys.exit(0 if ok else 1)
    if what == "bench":
        ok = run_bench_compare(prec)
        print("bench %s: %s (%.1f s)" % (prec, "BITWISE EQUAL" if ok else "DIFFERENT", time.time() - t0))
        sys.exit(0 if ok else 1)
    if what == "time":
        if len(sys.argv) > 3 and sys.argv[3] == "env":
            modes = (int(os.environ.get("SJ_TMA", "3")),)
        else:
            modes = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (0, 1, 2, 3)
        run_time(prec, modes)
