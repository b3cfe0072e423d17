"""Debug timeline of one time step (per-launch CUDA events): python scripts/trace_step.py [precision]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sim_juncs_b200.bound_geom import BoundGeom
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
st = bench.load_settings()
bg = BoundGeom(st, os.path.join(bench.ROOT, "scenes", "json", bench.SCENE + ".json"), precision=prec, n_sets=2)
bg.sim.run(40, 20)
L = bg.sim.L
L.sj_trace_step.argtypes = [ctypes.c_void_p]
for _ in range(2):
    print("---- step")
    sys.stdout.flush()
    L.sj_trace_step(bg.sim.h)
