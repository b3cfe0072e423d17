"""C4-sized sanity + throughput: Au_SiO2_box at --grid-res R (56 -> 1009^3 cells), python scripts/big_grid.py [res] [prec] [sets] [steps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sim_juncs_b200.bound_geom import BoundGeom
from sim_juncs_b200.settings import ParseSettings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res = float(sys.argv[1]) if len(sys.argv) > 1 else 56.0
prec = sys.argv[2] if len(sys.argv) > 2 else "f32"
sets = int(sys.argv[3]) if len(sys.argv) > 3 else 1
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
st = ParseSettings()
for k, v in json.load(open(os.path.join(ROOT, "scenes/json/Au_SiO2_box.json")))["settings"].items():
    setattr(st, k, v)
st.grid_num = -1; st.resolution = res; st.correct_defaults()
n = st.grid_cells()
t0 = time.time()
bg = BoundGeom(st, os.path.join(ROOT, "scenes/json/Au_SiO2_box.json"), precision=prec, n_sets=sets)
print("grid %d^3 = %.3g cells, rasterized + set up in %.2f s (raster %.3f s)" % (n, float(n) ** 3, time.time() - t0, bg.t_raster))
c = bg.sim.counts(); print(c)
bg.sim.run(900 if n < 300 else 10, 20)
ms = bg.sim.run_timed(steps, 20)
v = float(n) ** 3 * sets * steps / (ms * 1e-3)
print("%s x%d: %.3f ms/step, %.2f G cell-updates/s, step roofline frac %.3f" % (prec, sets, ms / steps, v / 1e9, bg.sim.bytes_per_step() / (ms * 1e-3 / steps) / 6548.2e9))
m = bg.sim.monitors()
print("monitors finite:", bool(np.isfinite(m).all()), "max", float(np.abs(m).max()))
ex = bg.sim.field(0, 0)
print("Ex finite:", bool(np.isfinite(ex).all()), "max", float(np.abs(ex).max()))
