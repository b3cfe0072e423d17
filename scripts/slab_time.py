"""Step time of ONE z-slab of the bench workload on one GPU, neighbours not connected (its boundary items neither wait nor
send): what the per-slab kernel of an N-GPU run costs without the exchange.  python scripts/slab_time.py kz0 kz1 [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sim_juncs_b200.bound_geom import BoundGeom
kz = (int(sys.argv[1]), int(sys.argv[2]))
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
st = bench.load_settings()
bg = BoundGeom(st, None, precision="f64", n_sets=2, kz=kz)
bg.sim.run(40, 20)
bg.sim.sync()
best = min(bg.sim.run_timed(steps, 20) for _ in range(3)) / steps
knobs = {k: v for k, v in os.environ.items() if k.startswith("SJ_")}
print("slab %s: ms/step %.4f  launches/step %.1f  %s" % (kz, best, bg.sim.launches() / (40 + 3 * steps), knobs))
