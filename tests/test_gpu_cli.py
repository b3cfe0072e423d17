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


# ---- the reference's OWN sim_geom on the B200 engine (host/meep_compat) -----------------------------------------
MEEP_EXE = os.path.join(ROOT, "host", "_ref", "sim_geom_meep")


@pytest.mark.skipif(not os.path.exists(MEEP_EXE), reason="host/_ref/sim_geom_meep is built where /root/reference exists")
@pytest.mark.parametrize("name", ["run_slabs", "cw_slab", "graphene_res2p5", "graphene_smooth2"])
def test_unmodified_reference_main_on_the_cuda_engine(name, tmp_path, golden):
    """main.cpp + disp.cpp + cgs*.cpp + data_utils.cpp of the reference, unmodified, linked against host/meep_compat
    (meep API slice on the C ABI) instead of libmeep."""
    _cli_against_reference_golden(MEEP_EXE, name, tmp_path, golden)


@pytest.mark.skipif(not os.path.exists(EXE), reason="host/_ref/sim_geom is built where /root/reference exists")
@pytest.mark.parametrize("name", ["cw_slab", "graphene_smooth2"])
def test_own_cpp_host_against_reference_driver(name, tmp_path, golden):
    """host/sim_geom (reference front end + GPU rasterizer, incl. smooth_n > 0) on the same launch lines."""
    _cli_against_reference_golden(EXE, name, tmp_path, golden)


def _cli_against_reference_golden(exe, name, tmp_path, golden):
    """A C++ binary on the engine, launched with the reference's command line: the series, spectra, locations, time bounds
    and source records it writes to field_samples.h5 must be the ones the reference's own sources produce over the CPU
    oracle (tests/golden/ref_*.npz), series and spectra to <= 1e-9."""
    from helpers import rel_l2
    from sim_juncs_b200 import hdf5
    g = np.load(os.path.join(golden, "ref_%s.npz" % name))
    argv = [exe, "--conf-file", str(g["conf"]), "--out-dir", str(tmp_path)] + [str(a) for a in g["argv"]]
    out = subprocess.check_output(argv, cwd=ROOT, timeout=900).decode()
    assert "finished writing hdf5 file!" in out
    f = hdf5.File(os.path.join(str(tmp_path), "field_samples.h5"))
    ref = g["time"]
    cols, fcols, locs = [], [], []
    for cn in sorted(k for k in f.keys() if k.startswith("cluster_")):
        cl = f[cn]
        loc = cl["locations"].read()
        if len(loc):
            locs.append(np.stack([loc["x"], loc["y"], loc["z"]], axis=1))
        for pn in sorted(k for k in cl.keys() if k.startswith("point_")):
            t = cl[pn]["time"].read()
            cols.append(t["Re"] + 1j * t["Im"])
            fr = cl[pn]["frequency"].read()
            fcols.append(fr["Re"] + 1j * fr["Im"])
    got = np.stack(cols, axis=1)
    assert got.shape == ref.shape
    assert np.abs(ref).max() > 1e-7
    assert rel_l2(got, ref) <= 1e-9, rel_l2(got, ref)
    assert rel_l2(np.stack(fcols, axis=1), g["frequency"]) <= 1e-9
    assert np.array_equal(np.concatenate(locs), g["locations"])
    assert np.array_equal(f["info"]["time_bounds"].read(), g["time_bounds"])
    assert int(f["info"]["n_time_points"].read()[0]) == int(g["n_time_points"][0])
    src = f["info"]["sources"].read()
    assert np.array_equal(np.stack([src[k] for k in ("wavelen", "width", "phase", "start_time", "end_time", "amplitude")], axis=1), g["sources"])
    assert sorted(f["info"]["cgs_params"].keys()) == sorted(str(x) for x in g["cgs_names"])


def test_custom_source_and_sample_at_equal_builtin_paths():
    """sj_add_custom_source with the Gaussian waveform evaluated in Python == sj_add_gaussian_source; sj_sample_at ==
    the monitor list, bit for bit."""
    import cmath
    import math
    from sim_juncs_b200 import Sim
    n, a = (24, 20, 28), 6.0
    lo, hi = [0, 0, 1.0], [n[0] / a, n[1] / a, 1.0]
    f0, w, ph, t0, t1 = 0.4, 1.5, 0.3, 1.0, 19.0
    omega, peak, cutoff = 2 * math.pi * f0, 0.5 * (t0 + t1), float(np.float32((t1 - t0) * 0.5))

    def dipole(t):                       # disp.cpp:395-400
        tt = t - peak
        if float(np.float32(abs(tt))) > cutoff:
            return 0.0
        return math.exp(-tt * tt / (2 * w * w)) * cmath.rect(1.0, -omega * tt - (ph + math.pi)) * (1.0 / complex(0, -omega))
    pts = np.array([[2.0, 1.6, 2.3], [1.0, 2.0, 3.0], [3.3, 0.2, 0.9]])
    sims = []
    for custom in (False, True):
        g = Sim(n, a, pml=1.0, n_sets=2)
        g.set_materials([(1.0, []), (2.25, [(1.1, 0.05, 1.3, 0)])], [(np.arange(29)[:, None, None] > 16).astype(np.uint8) * np.ones((29, 21, 25), dtype=np.uint8)] * 3)
        if custom:
            g.add_custom_source(0, lo, hi, 1.0, dipole, float(np.float32(peak + cutoff)), True)
        else:
            g.add_gaussian_source(0, lo, hi, 1.0, f0, w, ph, t0, t1, True)
        g.add_monitors(pts, 0)
        g.run(120, 1 << 30)
        sims.append(g)
    assert sims[0].last_source_time() == sims[1].last_source_time()
    a0, a1 = sims[0].sample_at(pts), sims[1].sample_at(pts)
    assert np.abs(a0).max() > 1e-6
    assert rel_l2_(a1, a0) < 1e-13              # libm exp / polar in C++ vs Python: last-bit differences of the waveform
    for g in sims:                              # the run loop samples BEFORE it steps (disp.cpp:719-741): this sample is of
        g.run(1, 1)                             # the state sample_at saw
    assert np.array_equal(sims[0].monitors()[-1], a0) and np.array_equal(sims[1].monitors()[-1], a1)


def rel_l2_(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


@pytest.mark.skipif(not os.path.exists(MEEP_EXE), reason="host/_ref/sim_geom_meep is built where /root/reference exists")
@pytest.mark.parametrize("exe", [MEEP_EXE, EXE])
def test_cpp_hosts_write_the_whole_grid_dumps(exe, tmp_path):
    """fields.output_hdf5 through host/meep_compat (the reference's main.cpp) and in the own C++ host: eps-000000.00.h5 at
    the start of run() and, with dump_raw = 1, one ex-<time>.h5 per save (disp.cpp:696, 732-737), against the same launch
    through the Python host (output.py)."""
    from sim_juncs_b200 import hdf5
    from sim_juncs_b200.settings import settings_from
    conf = "scenes/tests/run_dump.conf"
    a, b = str(tmp_path / "ref_main"), str(tmp_path / "py")
    os.makedirs(a)
    subprocess.check_output([exe, "--conf-file", conf, "--out-dir", a], cwd=ROOT, timeout=600)
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        st = settings_from(conf, ["--out-dir", b])
        bg = BoundGeom(st, None)
        bg.run(b)
    finally:
        os.chdir(cwd)
    fa = sorted(f for f in os.listdir(a) if f.endswith(".h5") and f != "field_samples.h5")
    fb = sorted(f for f in os.listdir(b) if f.endswith(".h5") and f != "field_samples.h5")
    assert fa == fb and "eps-000000.00.h5" in fa and len(fa) == 1 + (bg.n_t_pts + 39) // 40
    eps_a, eps_b = hdf5.File(os.path.join(a, "eps-000000.00.h5"))["eps"].read(), hdf5.File(os.path.join(b, "eps-000000.00.h5"))["eps"].read()
    n = st.grid_cells()
    assert eps_a.shape == (n, n, n) and np.array_equal(eps_a, eps_b) and eps_a.min() == 1.0 and eps_a.max() == 3.5
    seen = 0.0
    for f in fa[1:]:
        for ds in ("ex.r", "ex.i"):
            x, y = hdf5.File(os.path.join(a, f))[ds].read(), hdf5.File(os.path.join(b, f))[ds].read()
            assert x.shape == (n, n, n)
            assert np.abs(x - y).max() <= 1e-12 * max(np.abs(y).max(), 1e-30), (f, ds)
            seen = max(seen, np.abs(y).max())
    assert seen > 1e-9


@pytest.mark.skipif(not os.path.exists(EXE), reason="host/_ref/sim_geom is built where /root/reference exists")
def test_cpp_host_gpus_flag_equals_single_gpu(tmp_path):
    """host/sim_geom --gpus 3 (z-slabs inside the library; here all on GPU 0) writes the same field_samples.h5 series as
    the plain run"""
    from sim_juncs_b200 import hdf5
    conf = os.path.join(ROOT, "scenes", "tests", "graphene_short.conf")
    outs = []
    for tag, extra, env in (("one", [], {}), ("three", ["--gpus", "3"], {"SJ_ONE_DEVICE": "1"})):
        out = str(tmp_path / tag)
        os.makedirs(out)
        txt = subprocess.check_output([EXE, "--conf-file", conf, "--out-dir", out] + extra, cwd=ROOT, timeout=900,
                                      env=dict(os.environ, **env)).decode()
        assert "finished writing hdf5 file!" in txt
        if extra:
            assert "z-slabs:" in txt
        f = hdf5.File(os.path.join(out, "field_samples.h5"))
        cols = []
        for cn in sorted(k for k in f.keys() if k.startswith("cluster_")):
            for pn in sorted(k for k in f[cn].keys() if k.startswith("point_")):
                t = f[cn][pn]["time"].read()
                cols.append(t["Re"] + 1j * t["Im"])
        outs.append(np.stack(cols, axis=1))
    assert np.abs(outs[0]).max() > 1e-7
    assert np.array_equal(outs[0], outs[1])
