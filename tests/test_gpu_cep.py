"""sj_extract_cep (pulse parameters of every monitor series, on the device; csrc/sj_cep.cu) against the NumPy restatement
of the reference's phases.signal (oracle/cep_oracle.py, itself pinned to the reference class by tests/test_cep_oracle.py),
on the series of a real run, and timed against the per-series host loop the reference's scripts run."""
import time

import numpy as np
import pytest

from oracle import cep_oracle as co

pytestmark = pytest.mark.gpu


def test_device_extraction_equals_host_pipeline(scene_json):
    from helpers import early_pulse, settings_from_doc
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.scene import Scene
    path = scene_json("Au_graphene_box")
    st = settings_from_doc(path)
    st.grid_num = 73
    st.resolution = 73 / 18.0
    st.post_source_t = 12.0
    st.save_span = 2
    bg = BoundGeom(st, early_pulse(Scene.load(path), 0.6), n_sets=2, device=0)
    bg.run()
    ser = bg.sim.monitors()                      # [n_samples][n_mon][n_sets]
    n_s, n_mon, n_sets = ser.shape
    assert n_s > 200 and np.abs(ser).max() > 1e-4
    dt_s = bg.time_bounds()[2]                   # fs between two samples, as phases.py reads it from info/time_bounds
    t0 = time.perf_counter()
    dev = bg.sim.extract_cep(dt_s)
    t_dev = time.perf_counter() - t0
    t = np.arange(n_s) * dt_s
    t0 = time.perf_counter()
    checked = 0
    for q in range(n_sets):
        for m in range(n_mon):
            v = ser[:, m, q]
            if np.abs(v).max() < 1e-9 or dev["status"][q, m] & 4:
                continue
            ref = co.signal_params(t, v)
            assert dev["t0_ind"][q, m] == ref["t0_ind"], (q, m)
            assert dev["f0_ind"][q, m] == ref["f0_ind"], (q, m)
            assert dev["f_min"][q, m] == ref["f_min"] and dev["f_max"][q, m] == ref["f_max"], (q, m)
            assert abs(dev["f0"][q, m] - ref["f0"]) <= 1e-9 * abs(ref["f0"])
            assert abs(dev["t0_corr"][q, m] - ref["t0_corr"]) <= 1e-7 * max(1.0, abs(ref["t0_corr"]))
            d = abs(dev["phi_corr"][q, m] - ref["phi_corr"])
            assert min(d, 2 * np.pi - d) <= 1e-7, (q, m, dev["phi_corr"][q, m], ref["phi_corr"])
            checked += 1
    t_host = time.perf_counter() - t0
    assert checked >= n_mon
    print("cep: %d series of %d samples: device %.1f ms (one launch + one copy), host loop %.1f ms" %
          (n_mon * n_sets, n_s, 1e3 * t_dev, 1e3 * t_host))


def test_device_extraction_on_reference_goldens(golden):
    """the golden series of the reference class, pushed through the device kernel via a tiny simulation's series buffer is
    not possible without stepping; instead: synthetic pulses as sources of truth are covered by the CPU test, and the
    device path is compared with them through the restatement above.  Here: error paths."""
    from sim_juncs_b200 import Sim, SjError
    g = Sim((16, 16, 16), 4.0, n_sets=1, device=0)
    g.add_monitors([[1.0, 1.0, 1.0]], 0)
    g.run(4, 1)
    with pytest.raises(SjError):
        g.extract_cep(0.1)                       # fewer than 8 samples
