"""Which part of the Au_graphene_box scene makes the fields grow after ~6000 steps?  Full-length production runs (181^3,
31 159 steps) of variants of scenes/Au_graphene_box/junc.geom: python scripts/blowup_bisect.py [variant ...]"""
import os
import re
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sim_juncs_b200.bound_geom import BoundGeom  # noqa: E402
from sim_juncs_b200.settings import settings_from  # noqa: E402

SRC = open(os.path.join(ROOT, "scenes", "Au_graphene_box", "junc.geom")).read()
AU = re.search(r"//Au:.*?\n\]\)\n", SRC, re.S).group(0)
GR = re.search(r"//Graphene sheet.*?\n\]\)\n", SRC, re.S).group(0)
SI = re.search(r"//SiO2 substrate.*?\n\]\)\n", SRC, re.S).group(0)
AU_IN = AU.replace("Box([0, 0,    top], [length, left,   bot])", "Box([1.5, 1.5,  top], [length-1.5, left, bot])") \
          .replace("Box([0, rght, top], [length, length, bot])", "Box([1.5, rght, top], [length-1.5, length-1.5, bot])")
GR_IN = GR.replace("Box([0, left, sheet_lo], [length, rght, sheet_hi])", "Box([1.5, left, sheet_lo], [length-1.5, rght, sheet_hi])")
SI_IN = SI.replace("Box([0, 0, bot], [length, length, length])", "Box([1.5, 1.5, bot], [length-1.5, length-1.5, length-1.5])")
VARIANTS = {
    "original": SRC,
    "no_graphene": SRC.replace(GR, ""),
    "no_au": SRC.replace(AU, ""),
    "au_inside": SRC.replace(AU, AU_IN),
    "graphene_inside": SRC.replace(GR, GR_IN),
    "au_graphene_inside": SRC.replace(AU, AU_IN).replace(GR, GR_IN),
    "all_inside": SRC.replace(AU, AU_IN).replace(GR, GR_IN).replace(SI, SI_IN),
}


def run(name):
    d = tempfile.mkdtemp()
    open(os.path.join(d, "junc.geom"), "w").write(VARIANTS[name])
    st = settings_from(os.path.join(ROOT, "scenes", "Au_graphene_box", "params.conf"))
    st.geom_fname = os.path.join(d, "junc.geom")
    st.save_span = 20
    bg = BoundGeom(st, None)
    t0 = time.time()
    bg.run()
    m = np.abs(np.stack(bg.get_field_times(), axis=1)).max(axis=1)
    w = len(m) // 8
    print("%-20s %d steps in %.1f s  max|Ex| per eighth of the run: %s" % (name, bg.n_t_pts, time.time() - t0,
          " ".join("%.2e" % m[i * w:(i + 1) * w].max() for i in range(8))))
    sys.stdout.flush()


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(VARIANTS)):
        run(name)
