"""The CUDA path end to end (own CGS reader -> BoundGeom -> C ABI -> sm_100a kernels -> on-device spectra) against
golden series produced by the REFERENCE's own driver (main.cpp + disp.cpp + cgs*.cpp compiled in place over the
CPU oracle, scripts/make_ref_golden.py).  Same launch line on both sides.  Tolerances (north_star): monitor series
and spectra <= 1e-9 relative L2 in fp64, <= 1e-4 in fp32."""
import os

import numpy as np
import pytest

from helpers import ROOT, rel_l2
from sim_juncs_b200.bound_geom import BoundGeom
from sim_juncs_b200.settings import settings_from

pytestmark = pytest.mark.gpu
CASES = ["run_slabs", "cw_slab", "graphene_res2p5", "graphene_short", "run_slabs_smooth1", "graphene_smooth2"]


def _launch(name, golden, precision):
    g = np.load(os.path.join(golden, "ref_%s.npz" % name))
    cwd = os.getcwd()
    os.chdir(ROOT)                      # geom_fname in the conf files is relative to the repository root
    try:
        st = settings_from(str(g["conf"]), [str(a) for a in g["argv"]])
        bg = BoundGeom(st, None, precision=precision)
        bg.run()
    finally:
        os.chdir(cwd)
    return g, st, bg


@pytest.mark.parametrize("name", CASES)
def test_fp64_series_and_spectra_match_reference_driver(name, golden):
    g, st, bg = _launch(name, golden, "f64")
    ref = g["time"]
    n_saves = int(g["n_time_points"][0])
    assert bg.n_t_pts // bg.save_span == n_saves == ref.shape[0]
    got = np.stack(bg.get_field_times(), axis=1)[:n_saves]
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= 1e-9, rel_l2(got, ref)
    assert np.array_equal(np.array(bg.time_bounds()), g["time_bounds"])
    assert np.array_equal(np.array(bg.get_monitor_locs()), g["locations"])
    src = np.array([[s.wavelen, s.width, s.phase, s.start_time, s.end_time, s.amplitude] for s in bg.get_sources()])
    assert np.array_equal(src, g["sources"])
    # `frequency` as save_field_times writes it: the reference transforms ALL samples it pushed (one more than it writes
    # when save_span does not divide n_t_pts), so compare through the same path: output.field_samples_dict
    from sim_juncs_b200.output import field_samples_dict
    d = field_samples_dict(bg)
    fr = np.stack([d[k][:, 0] + 1j * d[k][:, 1] for k in d if k.endswith("/frequency")], axis=1)
    assert fr.shape == g["frequency"].shape
    assert rel_l2(fr, g["frequency"]) <= 1e-9, rel_l2(fr, g["frequency"])
    tm = np.stack([d[k][:, 0] + 1j * d[k][:, 1] for k in d if k.endswith("/time")], axis=1)
    assert tm.shape == ref.shape and rel_l2(tm, ref) <= 1e-9
    assert [k.rsplit("/", 1)[1] for k in d if k.startswith("info/cgs_params/")] == [str(x) for x in g["cgs_names"]]
    assert np.array_equal(np.array([d[k][0] for k in d if k.startswith("info/cgs_params/")]), g["cgs_values"])


@pytest.mark.parametrize("name", ["cw_slab", "graphene_res2p5"])
def test_fp32_series_match_reference_driver(name, golden):
    g, st, bg = _launch(name, golden, "f32")
    ref = g["time"]
    got = np.stack(bg.get_field_times(), axis=1)[:ref.shape[0]]
    assert rel_l2(got, ref) <= 1e-4, rel_l2(got, ref)


def test_smoothed_rasterizer_equals_reference_in_bound(golden):
    """smooth_n = 1: eps_inf at every Yee point from the CUDA rasterizer (counts over the point and its 8 offsets, material
    table from the distinct count tuples) against what the reference's own in_bound returned -- exact."""
    g = np.load(os.path.join(golden, "ref_run_slabs_smooth1.npz"))
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        st = settings_from(str(g["conf"]), [str(a) for a in g["argv"]])
        bg = BoundGeom(st, None, n_sets=1)
    finally:
        os.chdir(cwd)
    n = st.grid_cells()
    table = np.array([m[0] for m in bg.sim.material_table()])
    assert 2 < len(table) <= 10
    for c in range(3):
        eps = table[bg.sim.material_ids(c)]
        assert np.array_equal(eps, g["eps"][c].reshape(n + 1, n + 1, n + 1))
    # the plain inside bits are still reported
    assert set(np.unique(bg.sim.region_masks(0)).tolist()) == {0, 1}


def test_smoothed_slab_equals_global(golden):
    """z-slab rasterization with smoothing: material ids may be numbered differently per slab, eps may not differ."""
    g = np.load(os.path.join(golden, "ref_graphene_smooth2.npz"))
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        st = settings_from(str(g["conf"]), [str(a) for a in g["argv"]])
        whole = BoundGeom(st, None, n_sets=1)
        part = BoundGeom(st, None, n_sets=1, kz=(17, 33))
    finally:
        os.chdir(cwd)
    tw = np.array([m[0] for m in whole.sim.material_table()])
    tp = np.array([m[0] for m in part.sim.material_table()])
    for c in range(3):
        assert np.array_equal(tw[whole.sim.material_ids(c)][17:33], tp[part.sim.material_ids(c)])


def test_late_time_tail_matches_reference_driver(golden):
    """5965 steps at res 5 (the pulse is over after ~1100): the weak ringing that remains -- six orders of magnitude below
    the pulse -- must still be the reference driver's, window by window (absolute error against the pulse's scale, since
    round-off of the pulse is what limits the tail)."""
    g, st, bg = _launch("graphene_long", golden, "f64")
    ref = g["time"]
    got = np.stack(bg.get_field_times(), axis=1)[:ref.shape[0], ::int(g["monitor_stride"])]
    peak = np.abs(ref).max()
    assert rel_l2(got, ref) <= 1e-9
    for lo in range(0, ref.shape[0], 100):
        w = slice(lo, min(lo + 100, ref.shape[0]))
        assert np.abs(got[w] - ref[w]).max() <= 1e-12 * peak, (lo, np.abs(got[w] - ref[w]).max())
        assert rel_l2(got[w], ref[w]) <= 1e-6, (lo, rel_l2(got[w], ref[w]))
    assert np.abs(ref[300:]).max() < 1e-4 * peak            # no late-time growth at this resolution


def test_production_launch_of_the_shipped_Au_SiO2_box_scene(scene_json, golden):
    """SURVEY case P: the reference's own junctions/Au_SiO2_box scene as scripts/run.sh launches it (--grid-res 12 ->
    217^3 cells, 3506 steps, 1600 monitors, save_span 20).  Golden: the reference's driver over the CPU oracle (every
    20th monitor kept).  The GPU run enters through the scene fixture the reference parser produced for that launch."""
    from helpers import settings_from_doc
    if not os.path.exists(os.path.join(golden, "ref_Au_SiO2_box_P.npz")):
        pytest.skip("tests/golden/ref_Au_SiO2_box_P.npz not generated (scripts/make_ref_golden.py Au_SiO2_box_P, 20 min of CPU)")
    g = np.load(os.path.join(golden, "ref_Au_SiO2_box_P.npz"))
    st = settings_from_doc(scene_json("Au_SiO2_box"))
    bg = BoundGeom(st, scene_json("Au_SiO2_box"))
    bg.run()
    assert bg.sim.n[0] == 217 and bg.n_t_pts == 3506
    ref = g["time"]
    n_saves = int(g["n_time_points"][0])
    assert bg.n_t_pts // bg.save_span == n_saves == ref.shape[0] == 175
    got = np.stack(bg.get_field_times(), axis=1)
    assert got.shape[1] == 1600 and len(bg.monitor_clusters) == int(g["n_clusters"][0]) == 40
    got = got[:n_saves, ::int(g["monitor_stride"])]
    assert np.abs(ref).max() > 1e-3
    assert rel_l2(got, ref) <= 1e-9, rel_l2(got, ref)
    assert np.array_equal(np.array(bg.time_bounds()), g["time_bounds"])
    assert np.array_equal(np.array(bg.get_monitor_locs()), g["locations"])
    from sim_juncs_b200.output import field_samples_dict
    d = field_samples_dict(bg)
    fr = np.stack([d[k][:, 0] + 1j * d[k][:, 1] for k in d if k.endswith("/frequency")], axis=1)[:, ::int(g["monitor_stride"])]
    assert fr.shape == g["frequency"].shape == (128, ref.shape[1])
    assert rel_l2(fr, g["frequency"]) <= 1e-9


def test_shipped_bowtie_scene_on_a_coarse_grid(scene_json, golden):
    """BASELINE config 3: the reference's junctions/Au_SiO2_bowtie scene (rotated boxes, cylinders, 1600 monitors) as
    `--grid-res 4 --opts "width=0.05;..."` launches it (73^3 cells).  Golden: the reference's driver over the CPU oracle,
    every 20th monitor; the GPU run enters through the reference parser's fixture with the same grid."""
    from helpers import settings_from_doc
    g = np.load(os.path.join(golden, "ref_bowtie_res4.npz"))
    st = settings_from_doc(scene_json("Au_SiO2_bowtie"))
    st.grid_num = 73
    st.resolution = 73 / 18.0          # argparse.h:171-191 for --grid-res 4.0: odd grid number over the total length
    bg = BoundGeom(st, scene_json("Au_SiO2_bowtie"))
    bg.run()
    ref = g["time"]
    n_saves = int(g["n_time_points"][0])
    assert bg.sim.n[0] == 73 and bg.n_t_pts // bg.save_span == n_saves == ref.shape[0]
    got = np.stack(bg.get_field_times(), axis=1)
    assert got.shape[1] == 1600
    got = got[:n_saves, ::int(g["monitor_stride"])]
    assert np.abs(ref).max() > 1e-3
    assert rel_l2(got, ref) <= 1e-9, rel_l2(got, ref)
    assert np.array_equal(np.array(bg.time_bounds()), g["time_bounds"])
    assert np.array_equal(np.array(bg.get_monitor_locs()), g["locations"])


def test_phase_batch_against_reference_driver(golden):
    """BASELINE config 5: quartz_box run as a CEP phase batch (one real field set per phase, shared materials).  The sets
    for phase 0 and pi/2 are the real and imaginary parts of the complex run -- here checked against the series the
    reference's own driver produced for that scene (complex fields, as meep runs them)."""
    g = np.load(os.path.join(golden, "ref_quartz_res6.npz"))
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        st = settings_from(str(g["conf"]), [str(a) for a in g["argv"]])
        bg = BoundGeom(st, None, phases=[0.0, np.pi / 2, 0.7])
        bg.run()
    finally:
        os.chdir(cwd)
    ref = g["time"]
    n = ref.shape[0]
    per_phase = [np.stack(ft, axis=1)[:n] for ft in bg.get_field_times()]       # [phase][save, monitor]
    assert np.abs(ref).max() > 1e-4
    assert rel_l2(per_phase[0].real, ref.real) <= 1e-9, rel_l2(per_phase[0].real, ref.real)
    assert rel_l2(per_phase[1].real, ref.imag) <= 1e-9, rel_l2(per_phase[1].real, ref.imag)
    # any other phase is the rotation of the complex series: Re(E e^{-i phi})
    want = (ref * np.exp(-1j * 0.7)).real
    assert rel_l2(per_phase[2].real, want) <= 1e-9, rel_l2(per_phase[2].real, want)
