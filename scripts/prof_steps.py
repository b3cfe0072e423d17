"""Short run of the bench workload for ncu captures: python scripts/prof_steps.py [steps] [precision]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sim_juncs_b200.bound_geom import BoundGeom
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
prec = sys.argv[2] if len(sys.argv) > 2 else "f64"
st = bench.load_settings()
bg = BoundGeom(st, os.path.join(bench.ROOT, "scenes", "json", bench.SCENE + ".json"), precision=prec, n_sets=2)
bg.sim.run(steps, 20)
print("done", bg.sim.launches())
