import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import rel_l2
from sim_juncs_b200.bound_geom import BoundGeom
from sim_juncs_b200.settings import settings_from
for name in ("cw_slab", "graphene_res2p5", "graphene_short", "graphene_smooth2"):
    g = np.load("tests/golden/ref_%s.npz" % name)
    st = settings_from(str(g["conf"]), [str(a) for a in g["argv"]])
    for prec in ("f32", "f64"):
        bg = BoundGeom(st, None, precision=prec); bg.run()
        got = np.stack(bg.get_field_times(), axis=1)[:g["time"].shape[0]]
        print(name, prec, "%.3e" % rel_l2(got, g["time"]))
