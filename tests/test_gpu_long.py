"""The target configuration run to completion (verdict r1: the number in BENCH must not be the first steps of a run that
overflows).  scenes/Au_graphene_box/params_inside.conf = BASELINE configs[1] at production resolution (181^3 cells, 2
field sets, fp64, 31 159 steps = 300 fs after the pulse) with the Au leads and the graphene sheet ended 0.5 units before
the absorbing layer (junc_inside.geom; a Drude metal inside the UPML is what makes junc.geom grow, scripts/
blowup_bisect.py).  About 17 s on one B200."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_full_production_run_stays_bounded(root):
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.settings import settings_from
    d = os.path.join(root, "scenes", "Au_graphene_box")
    st = settings_from(os.path.join(d, "params_inside.conf"))
    st.geom_fname = os.path.join(d, "junc_inside.geom")
    st.save_span = 20
    bg = BoundGeom(st, None)
    assert st.grid_cells() == 181
    bg.run()
    assert bg.n_t_pts > 31000
    m = np.abs(np.stack(bg.get_field_times(), axis=1)).max(axis=1)
    assert np.isfinite(m).all()
    w = len(m) // 8
    eighths = [m[i * w:(i + 1) * w].max() for i in range(8)]
    assert eighths[0] > 0.1                                   # the pulse
    assert all(b < a for a, b in zip(eighths[1:], eighths[2:]))   # ring-down: every later eighth is quieter
    assert eighths[-1] < 1e-5 * eighths[0]
