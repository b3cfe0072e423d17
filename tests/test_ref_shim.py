"""The reference's OWN driver -- src/main.cpp + src/disp.cpp + cgs*.cpp + data_utils.cpp compiled in place against
oracle/shim/ (a meep-API slice over oracle/fdtd_oracle.c and an HDF5 recorder) -- against this repository's host
logic.  Everything outside meep's arithmetic is the reference's real code there: scene parsing and context seeding,
structure_from_settings, cgs_material_function::in_bound, source_info, gaussian_src_time_phase::dipole, the run loop,
fft and save_field_times.  These tests pin the Python host (settings, cgs, scene, BoundGeom's unit conversions, the
oracle front end the GPU parity tests use, output.py) against it.  CPU only."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import ROOT, orc
from sim_juncs_b200.scene import Scene
from sim_juncs_b200.settings import ParseSettings

pytestmark = pytest.mark.skipif(not helpers.have_ref_sim_geom(), reason="oracle/_ref/sim_geom_ref not built and no /root/reference")


def _settings(conf, extra=()):
    from sim_juncs_b200.settings import settings_from
    return settings_from(conf, list(extra))


def _python_host_on_oracle(conf, extra=()):
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        st = _settings(conf, extra)
        scene = Scene.from_geom(st.geom_fname, st)
        masks = [helpers.oracle_raster(scene, st, c) for c in range(3)]
        o, n_t_pts = helpers.oracle_bound_geom(scene, st, masks)       # smooth_n > 0: counts instead of the masks
        o.run(n_t_pts, st.save_span or 1)
    finally:
        os.chdir(cwd)
    ser = o.monitors()
    return st, scene, o, n_t_pts, ser[:, :, 0] + 1j * ser[:, :, 1]


RUN_SH_OPTS = "width=0.05;thick=0.2;inf_thick=0;wavelen=0.76;n_cycles=0.5"       # reference scripts/run.sh:76-78
BOWTIE = ("/root/reference/junctions/Au_SiO2_bowtie/params.conf",
          ("--geom-file", "/root/reference/junctions/Au_SiO2_bowtie/junc.geom", "--grid-res", "4.0", "--opts", RUN_SH_OPTS))


def test_shipped_bowtie_scene_reference_driver_equals_python_host(tmp_path):
    """BASELINE config 3: the reference's own junctions/Au_SiO2_bowtie scene (rotated boxes, cylinders, --opts overrides,
    1600 monitors) on a coarse grid (--grid-res 4 -> 73^3): the reference's driver against settings.py + cgs.py + scene.py
    over the same engine, eps, sigma and series bit for bit."""
    if not os.path.exists(BOWTIE[0]):
        pytest.skip("needs /root/reference")
    test_reference_driver_equals_python_host(BOWTIE[0], BOWTIE[1], tmp_path)


def test_reference_own_test_scene_c1(tmp_path):
    """BASELINE config 0 literally: the reference's tests/run.conf + tests/run.geom (brace-form geometry = a root without
    children = vacuum) with the sizes its unit test uses (main_test.cpp:1884-1899: length 2, resolution 5)."""
    conf = "/root/reference/tests/run.conf"
    if not os.path.exists(conf):
        pytest.skip("needs /root/reference")
    test_reference_driver_equals_python_host(conf, ("--geom-file", "/root/reference/tests/run.geom", "--length", "2.0",
                                                    "--grid-res", "5.0"), tmp_path)


@pytest.mark.parametrize("conf,extra", [("scenes/tests/run.conf", ()), ("scenes/tests/cw_slab.conf", ()),
                                        ("scenes/tests/graphene_short.conf", ("--grid-res", "2.5")),
                                        ("scenes/tests/run_smooth.conf", ()), ("scenes/tests/graphene_smooth.conf", ())])
def test_reference_driver_equals_python_host(conf, extra, tmp_path):
    """Vacuum + dielectric slabs with a Gaussian pulse; a CW source; the Au / graphene / SiO2 junction (Drude + Lorentz
    poles, make_2d sheet) on a coarse grid; the last two with stochastic boundary smoothing (smooth_n = 1 and 2)."""
    entries, blob, out = helpers.run_ref_sim_geom(conf, str(tmp_path), extra)
    st, scene, o, n_t_pts, series = _python_host_on_oracle(conf, extra)
    # eps_inf at every Yee point: the reference's in_bound() vs region masks + material table
    n = st.grid_cells()
    for c, nm in enumerate("xyz"):
        eps_ref = np.fromfile(os.path.join(str(tmp_path), "eps_%s.f64" % nm)).reshape(n + 1, n + 1, n + 1)
        assert np.array_equal(1.0 / eps_ref, o.field("chi1inv", c))
    # sigma of every susceptibility (one per region and pole, disp.cpp:529-548) vs sigma_rp * inside bit
    amb, reps, rpoles = helpers.region_tables(scene, st)
    isus = 0
    most_levels = 0
    for r, poles in enumerate(rpoles):
        for (w0, g, sg, drude) in poles:
            for c, nm in enumerate("xyz"):
                sig_ref = np.fromfile(os.path.join(str(tmp_path), "sigma_%d_%s.f64" % (isus, nm))).reshape(n + 1, n + 1, n + 1)
                if st.smooth_n > 0:
                    cnt, tot = helpers.oracle_raster_counts(scene, st, c)
                    assert np.array_equal(sig_ref, 0.0 + (sg - 0.0) * cnt[r].astype(np.float64) / (tot + 1))
                    most_levels = max(most_levels, len(np.unique(sig_ref)))
                else:
                    bit = (helpers.oracle_raster(scene, st, c) >> r) & 1
                    assert np.array_equal(sig_ref, sg * bit)
            isus += 1
    assert not os.path.exists(os.path.join(str(tmp_path), "sigma_%d_x.f64" % isus))
    assert most_levels > 2 or st.smooth_n == 0 or isus == 0        # smoothing really produced intermediate sigma levels
    # the run loop: number of steps and saves, then the monitor series -- same engine underneath, so the only
    # differences could come from the host logic (waveform, placement box, amplitude, units, cadence)
    ref = helpers.ref_series(entries, blob)
    save_span = st.save_span or 1
    assert int(helpers.ref_dataset(entries, blob, "/info/n_time_points")[0]) == n_t_pts // save_span
    # sic: the reference reserves n_t_pts/save_span (+1) but pushes ceil(n_t_pts/save_span) samples and writes the first
    # n_t_pts/save_span of them
    assert ref.shape[0] == n_t_pts // save_span and ref.shape[1] == len(scene.monitor_locs)
    assert np.abs(ref).max() > 1e-6
    assert np.array_equal(ref, series[:ref.shape[0]])


@pytest.mark.parametrize("conf,extra", [("scenes/tests/run.conf", ()), ("scenes/tests/graphene_short.conf", ("--grid-res", "2.5"))])
def test_reference_field_samples_layout_equals_own_writer(conf, extra, tmp_path):
    """Every group, dataset, shape, compound member (name, offset, size) and value the reference's save_field_times
    hands to HDF5, against sim_juncs_b200/output.py + hdf5.py."""
    from sim_juncs_b200 import hdf5
    from sim_juncs_b200.output import save_field_samples
    entries, blob, out = helpers.run_ref_sim_geom(conf, str(tmp_path / "ref"), extra)
    st, scene, o, n_t_pts, series = _python_host_on_oracle(conf, extra)

    class Bg:        # the attributes save_field_samples reads from a BoundGeom
        pass
    bg = Bg()
    bg.problem, bg.settings, bg.um_scale, bg.save_span = scene, st, st.um_scale, st.save_span or 1
    bg.n_t_pts, bg.sources, bg.n_sets, bg.phases = n_t_pts, list(scene.sources), 2, None
    bg.monitor_locs = [tuple(p) for p in scene.monitor_locs]
    bg.monitor_clusters = list(scene.monitor_clusters)
    bg.field_times = [series[:, j].copy() for j in range(series.shape[1])]
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.scene import LIGHT_SPEED
    bg.ttot = o.last_source_time() + st.post_source_t * LIGHT_SPEED * st.um_scale
    bg.meep_time_to_fs = lambda t: BoundGeom.meep_time_to_fs(bg, t)
    bg.time_bounds = lambda: BoundGeom.time_bounds(bg)
    path = save_field_samples(bg, str(tmp_path / "own"))
    f = hdf5.File(path, "r")
    want_groups = [e["path"] for e in entries if e["what"] == "group"]
    want_sets = [e for e in entries if e["what"] == "dataset"]
    got_groups, got_sets = [], {}

    def walk(g, prefix):
        for k in g.keys():
            obj = g[k]
            p = prefix + "/" + k
            if hasattr(obj, "keys"):
                got_groups.append(p)
                walk(obj, p)
            else:
                got_sets[p] = obj
    walk(f, "")
    assert sorted(got_groups) == sorted(want_groups)
    assert sorted(got_sets) == sorted(e["path"] for e in want_sets)
    for e in want_sets:
        ds = got_sets[e["path"]]
        arr = np.asarray(ds[...] if hasattr(ds, "__getitem__") else ds)
        assert list(arr.shape[:1]) == e["dims"], e["path"]
        t = e["type"]
        ref = helpers.ref_dataset(entries, blob, e["path"])
        if t["kind"] == "compound":
            assert arr.dtype.names == tuple(m["name"] for m in t["members"]), e["path"]
            assert arr.dtype.itemsize == t["size"], e["path"]
            assert [arr.dtype.fields[m["name"]][1] for m in t["members"]] == [m["offset"] for m in t["members"]], e["path"]
            got = np.stack([arr[m["name"]] for m in t["members"]], axis=1) if len(arr) else np.zeros((0, len(t["members"])))
            if e["path"].endswith("/frequency"):
                assert np.allclose(got, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max())), e["path"]
            else:
                assert np.array_equal(got, ref), e["path"]
        else:
            assert arr.dtype == (np.dtype("<u8") if t["kind"] == "u64" else np.dtype("<f8")), e["path"]
            assert np.array_equal(arr, ref), e["path"]
