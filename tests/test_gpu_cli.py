"""The C++ host (host/sim_geom.cpp: the reference's own argparse.h + CGS parser compiled in place, on
the C ABI) against the Python mirror: same conf / geom files in, identical monitor series out."""
import os
import subprocess

import numpy as np
import pytest

from sim_juncs_b200.bound_geom import BoundGeom
from sim_juncs_b200.scene import Scene
from sim_juncs_b200.settings import ParseSettings

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "_ref", "sim_geom")


@pytest.mark.skipif(not os.path.exists(EXE), reason="host/_ref/sim_geom is built only where /root/reference exists")
def test_cpp_cli_matches_python_mirror(tmp_path, scene_json):
    conf = os.path.join(ROOT, "scenes", "tests", "graphene_short.conf")
    out = subprocess.check_output([EXE, "--conf-file", conf, "--out-dir", str(tmp_path)], cwd=ROOT, timeout=600).decode()
    assert "simulation completed in" in out and "initialization completed in" in out
    z = np.load(os.path.join(str(tmp_path), "field_samples.npz"))
    st = ParseSettings()
    st.parse_args(["--conf-file", conf, "--out-dir", str(tmp_path)])
    st.parse_conf_file(conf)
    st.correct_defaults()
    assert st.grid_cells() == 91
    bg = BoundGeom(st, scene_json("Au_graphene_box"), n_sets=2)     # same scene through the JSON fixture
    bg.run()
    assert int(z["info/n_clusters"][0]) == 1 and z["cluster_0/locations"].shape == (50, 3)
    assert np.allclose(z["info/time_bounds"], bg.time_bounds(), rtol=1e-13)
    assert np.array_equal(z["cluster_0/locations"], np.array(bg.get_monitor_locs()))
    assert abs(z["info/cgs_params/res"][0] - 181 / 18) < 1e-12
    n_saves = bg.n_t_pts // st.save_span
    worst = 0.0
    for i, series in enumerate(bg.get_field_times()):
        t = z["cluster_0/point_%02d/time" % i]
        got = t[:, 0] + 1j * t[:, 1]
        assert len(got) == n_saves          # sic: ceil(n_t_pts / save_span) sampled, n_t_pts / save_span written (disp.cpp:771)
        worst = max(worst, float(np.abs(got - series[:n_saves]).max()))
    assert worst == 0.0
    assert max(np.abs(s_).max() for s_ in bg.get_field_times()) > 1e-6
    # field_samples.h5 of the C++ host (host/sj_hdf5.hpp) against the Python mirror's own file, dataset by dataset
    from sim_juncs_b200 import hdf5
    from sim_juncs_b200.output import reference_fft
    mine = hdf5.File(bg.save_field_times(str(tmp_path / "py")))
    cpp = hdf5.File(os.path.join(str(tmp_path), "field_samples.h5"))
    assert cpp.keys() == mine.keys() and cpp["cluster_0"].keys() == mine["cluster_0"].keys()
    assert cpp["info"]["cgs_params"].keys() == mine["info"]["cgs_params"].keys()
    sa, sb = cpp["info"]["sources"].read(), mine["info"]["sources"].read()
    assert sa.dtype == sb.dtype and all(np.array_equal(sa[f], sb[f]) for f in sa.dtype.names)
    assert cpp["info"]["n_time_points"][0] == mine["info"]["n_time_points"][0]
    for pt in ("point_00", "point_49"):
        a, b = cpp["cluster_0"][pt]["time"].read(), mine["cluster_0"][pt]["time"].read()
        assert np.array_equal(a["Re"][:len(b)], b["Re"]) and np.array_equal(a["Im"][:len(b)], b["Im"])
        fa = cpp["cluster_0"][pt]["frequency"].read()
        fb = reference_fft(bg.get_field_times()[int(pt[-2:])])       # the transform takes every sample pushed, not only the written ones
        assert len(fa) == len(fb) and np.allclose(fa["Re"] + 1j * fa["Im"], fb, rtol=1e-9, atol=1e-12 * np.abs(fb).max())


def test_python_cli_with_own_parser_matches_reference_parser_path(tmp_path, scene_json):
    """`python -m sim_juncs_b200` reads the .geom with this repo's own CGS reader; the run must be
    bit-identical to the same scene entering through the reference parser's JSON dump."""
    import sys
    conf = os.path.join(ROOT, "scenes", "tests", "graphene_short.conf")
    out = subprocess.check_output([sys.executable, "-m", "sim_juncs_b200", "--conf-file", conf, "--out-dir", str(tmp_path)],
                                  cwd=ROOT, timeout=600).decode()
    assert "simulation completed in" in out and "saving timeseries completed in" in out
    z = np.load(os.path.join(str(tmp_path), "field_samples.npz"))
    st = ParseSettings()
    st.parse_args(["--conf-file", conf, "--out-dir", str(tmp_path)])
    st.parse_conf_file(conf)
    st.correct_defaults()
    bg = BoundGeom(st, scene_json("Au_graphene_box"), n_sets=2)
    bg.run()
    assert np.array_equal(z["cluster_0/locations"], np.array(bg.get_monitor_locs()))
    worst = 0.0
    for i, series in enumerate(bg.get_field_times()):
        t = z["cluster_0/point_%02d/time" % i]
        got = t[:, 0] + 1j * t[:, 1]
        worst = max(worst, float(np.abs(got - series[:len(got)]).max()))
    assert worst == 0.0
    assert max(np.abs(s_).max() for s_ in bg.get_field_times()) > 1e-6
