"""Diagnostic: GPU vs oracle on a reduced-grid junction scene for several pulse placements.
python scripts/dev_reduced.py [scene] [grid]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import oracle_bound_geom, settings_from_doc
from sim_juncs_b200.bound_geom import BoundGeom
from sim_juncs_b200.scene import Scene
name = sys.argv[1] if len(sys.argv) > 1 else "Au_SiO2_box"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 91
path = os.path.join(ROOT, "scenes", "json", name + ".json")
for start, width, steps in ((0.0, 0.3, 420), (0.5, 0.3, 480), (0.0, 0.3, 60), (0.0, 0.3, 150), (0.5, 0.3, 150)):
    st = settings_from_doc(path)
    st.grid_num = grid
    st.resolution = grid / (2 * (st.len / 2 + st.pml_thickness))
    sc = Scene.load(path)
    src = sc.sources[0]
    src.start_time, src.width, src.end_time = start, width, start + 12.0 * width
    bg = BoundGeom(st, sc, n_sets=2)
    masks = [bg.sim.region_masks(c) for c in range(3)]
    o, _ = oracle_bound_geom(sc, st, masks, nsets=2)
    o.run(steps, 10)
    bg.sim.run(steps, 10)
    out = []
    for kind, off in (("E", 0), ("H", 3)):
        for q in range(2):
            scale = max(np.linalg.norm(o.field(kind, c, q)) for c in range(3))
            for c in range(3):
                d = bg.sim.field(off + c, q) - o.field(kind, c, q)
                w = np.unravel_index(np.abs(d).argmax(), d.shape)
                out.append((np.linalg.norm(d) / scale, kind, c, q, w, float(np.abs(d).max()), scale))
    worst = max(out)
    print("start %.2f width %.2f steps %d: worst rel %.3e (%s%d set %d, max |d| %.3e at k,j,i=%s, scale %.3e)" % (
        start, width, steps, worst[0], worst[1], worst[2], worst[3], worst[5], worst[4], worst[6]))
