"""Step time of the bench workload without the rest of bench.py: python scripts/quick_time.py [steps] [precision]
(environment knobs of csrc/sj_engine.cu apply: SJ_ZC_FACE, SJ_ZC_GEN, SJ_ZCHUNK, SJ_N_AUX, ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sim_juncs_b200.bound_geom import BoundGeom
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
prec = sys.argv[2] if len(sys.argv) > 2 else "f64"
st = bench.load_settings()
bg = BoundGeom(st, os.path.join(bench.ROOT, "scenes", "json", bench.SCENE + ".json"), precision=prec, n_sets=2)
bg.sim.run(40, 20)
bg.sim.sync()
best = min(bg.sim.run_timed(steps, 20) for _ in range(3)) / steps
knobs = {k: v for k, v in os.environ.items() if k.startswith("SJ_")}
print("ms/step %.4f  launches/step %.1f  %s" % (best, bg.sim.launches() / (40 + 3 * steps), knobs))
