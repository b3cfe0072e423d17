"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerances (north_star): bit-exact for the voxelization (integer masks); monitor time series
<= 1e-9 relative L2 in fp64 and <= 1e-4 in fp32.  The fp64 engine is in practice at 1e-13 of the
oracle, so the fp64 bound asserted here is 1e-11."""
import numpy as np
import pytest

from helpers import early_pulse, oracle_bound_geom, region_tables, rel_l2, settings_from_doc
from oracle.oracle import OracleSim
from sim_juncs_b200 import Sim
from sim_juncs_b200.bound_geom import BoundGeom
from sim_juncs_b200.materials import materials_from_regions
from sim_juncs_b200.scene import LIGHT_SPEED, Scene

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-11, "f32": 1e-4}


# ---------------------------------------------------------------- voxelization: bit exact
@pytest.mark.parametrize("name", ["tests_run_slabs", "Au_SiO2_box", "Au_SiO2_bowtie", "Au_graphene_box"])
def test_rasterizer_bit_exact(name, scene_json, golden):
    g = np.load(golden + "/masks_%s.npz" % name)
    st = settings_from_doc(scene_json(name))
    bg = BoundGeom(st, scene_json(name), n_sets=1)
    for comp, key in enumerate(("ex", "ey", "ez")):
        got = bg.sim.region_masks(comp)
        assert np.array_equal(got, g[key]), "%s: %d points differ" % (key, int((got != g[key]).sum()))
    # material table follows in_bound(): ambient counted once per region
    nreg = len(bg.problem.regions)
    tab = bg.sim.material_table()
    assert len(tab) == 1 << nreg and tab[0][0] == st.ambient_eps * max(nreg, 1)


def test_rasterizer_slab_equals_global(scene_json, golden):
    g = np.load(golden + "/masks_Au_graphene_box.npz")
    st = settings_from_doc(scene_json("Au_graphene_box"))
    bg = BoundGeom(st, scene_json("Au_graphene_box"), n_sets=1, kz=(60, 121))
    for comp, key in enumerate(("ex", "ey", "ez")):
        assert np.array_equal(bg.sim.region_masks(comp), g[key][60:121])


# ---------------------------------------------------------------- stepping vs oracle
def _masks(n, a):
    shape = (n[2] + 1, n[1] + 1, n[0] + 1)
    k, j, i = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    ms = []
    for c in range(3):
        x = (i + 0.5 * (c == 0)) / a
        z = (k + 0.5 * (c == 2)) / a
        r0 = (z > 0.55 * n[2] / a)
        r1 = (x < 0.45 * n[0] / a) & (z > 0.3 * n[2] / a) & (z <= 0.62 * n[2] / a)
        ms.append(r0.astype(np.uint8) | (r1.astype(np.uint8) << 1))
    return ms


def _run_pair(n, a, pml, nsets, comp, lo, hi, mats, steps, prec, integrated=True, span=1, probe_step=None, cw=False):
    o = OracleSim(n, a, pml=pml, nsets=nsets)
    g = Sim(n, a, pml=pml, n_sets=nsets, precision=prec)
    if mats is not None:
        amb, reps, rpoles, masks = mats
        o.set_regions(amb, reps, rpoles, masks)
        g.set_materials(materials_from_regions(amb, reps, rpoles), masks)
    if cw:      # CW_source -> meep::continuous_src_time(f, width, start, end), slowness 3 (disp.cpp:615-619)
        src = (1.0, 0.4, 1.2, 0.5, 0.75 * steps * 0.5 / a, 3.0, integrated)
        o.add_cw_source(comp, lo, hi, *src)
        g.add_cw_source(comp, lo, hi, *src)
    else:
        src = (1.0, 0.4, 1.5, 0.3, 1.0, 1.0 + 12 * 1.5, integrated)
        o.add_gaussian_source(comp, lo, hi, *src)
        g.add_gaussian_source(comp, lo, hi, *src)
    L = [x / a for x in n]
    mon = [[L[0] / 2, L[1] / 2, L[2] / 2], [L[0] * 0.31, L[1] * 0.77, L[2] * 0.6], [L[0] / 3, L[1], L[2] / 3],
           [0.01, 0.02, 0.03], [L[0] * 0.5, L[1] * 0.5, 1.0]]
    o.add_monitors(mon, comp)
    g.add_monitors(mon, comp)
    probe_step = probe_step or steps
    o.run(probe_step, span)
    g.run(probe_step, span)
    # whole-field comparison while the pulse is inside the box, each component normalised by the
    # largest component of its kind (symmetry-forbidden components are pure round-off)
    ferr = 0.0
    for kind, off in (("E", 0), ("H", 3)):
        for q in range(nsets):
            scale = max(np.linalg.norm(o.field(kind, c, q)) for c in range(3))
            for c in range(3):
                ferr = max(ferr, np.linalg.norm(g.field(off + c, q) - o.field(kind, c, q)) / scale)
    if steps > probe_step:
        o.run(steps - probe_step, span)
        g.run(steps - probe_step, span)
    mo, mg = o.monitors()[:, :, :nsets], g.monitors()
    assert mo.shape == mg.shape and np.abs(mo).max() > 1e-3
    return rel_l2(mg, mo), ferr


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("integrated", [True, False])
def test_vacuum_pml_c1(prec, integrated):
    # config C1 (tests/run at the unit-test override): 20^3, a = 5, pml 1, Ey plane at z = 1, complex fields
    merr, ferr = _run_pair((20, 20, 20), 5.0, 1.0, 2, 1, [0, 0, 1], [4, 4, 1], None, 234, prec, integrated, probe_step=80)
    assert merr < TOL[prec] and ferr < TOL[prec]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_dispersive_pml_noncubic(prec):
    n, a = (44, 36, 40), 8.0
    lor, dru = (1.1, 0.05, 1.3, 0), (1e-10, 0.04, 2.0e19, 1)
    mats = (1.0, [2.25, 1.0], [[lor], [dru, (0.9, 0.2, 0.7, 0)]], _masks(n, a))
    merr, ferr = _run_pair(n, a, 1.0, 2, 0, [0, 0, 1.0], [n[0] / a, n[1] / a, 1.0], mats, 400, prec, span=3, probe_step=150)
    assert merr < TOL[prec] and ferr < TOL[prec]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_dispersive_no_pml_volume_source(prec):
    n, a = (44, 36, 40), 8.0
    mats = (1.0, [2.25, 1.0], [[(1.1, 0.05, 1.3, 0)], [(1e-10, 0.04, 2.0e19, 1)]], _masks(n, a))
    merr, ferr = _run_pair(n, a, 0.0, 1, 2, [1.0, 1.0, 1.0], [2.0, 3.0, 2.2], mats, 150, prec)
    assert merr < TOL[prec] and ferr < TOL[prec]


@pytest.mark.parametrize("integrated", [True, False])
def test_cw_source(integrated):
    """CW_source: tanh turn-on, steady oscillation, turn-off at 3/4 of the run, then ring-down into the PML."""
    n, a = (28, 24, 32), 6.0
    mats = (1.0, [2.25, 1.0], [[(1.1, 0.05, 1.3, 0)], []], _masks(n, a))
    merr, ferr = _run_pair(n, a, 1.0, 2, 0, [0, 0, 1.2], [n[0] / a, n[1] / a, 1.2], mats, 360, "f64", integrated,
                           probe_step=200, cw=True)
    assert merr < 1e-11 and ferr < 1e-11


def test_ragged_grid_sizes():
    # odd / non-multiple-of-anything extents exercise every tile edge
    for n in [(21, 23, 25), (33, 20, 47), (70, 21, 22)]:
        merr, ferr = _run_pair(n, 6.0, 1.0, 1, 0, [0, 0, 1.2], [n[0] / 6.0, n[1] / 6.0, 1.2], None, 120, "f64", probe_step=60)
        assert merr < 1e-11 and ferr < 1e-11, n


# ---------------------------------------------------------------- bound_geom level, reference configs
def _bound_geom_pair(name, scene_json, prec, max_steps=None):
    st = settings_from_doc(scene_json(name))
    sc = Scene.load(scene_json(name))
    bg = BoundGeom(st, scene_json(name), precision=prec, n_sets=2)
    masks = [bg.sim.region_masks(c) for c in range(3)]       # bit-exactness is asserted separately
    o, n_t = oracle_bound_geom(sc, st, masks, nsets=2)
    dt = bg.sim.dt
    assert int((bg.ttot + dt / 2) / dt) == n_t
    if max_steps:
        n_t = min(n_t, max_steps)
        bg.ttot = (n_t - 0.25) * dt
    bg.run()
    assert bg.n_t_pts == n_t
    o.run(n_t, st.save_span)
    series = np.array(bg.get_field_times()).T                  # [saves][mon]
    mo = o.monitors()
    return series, mo[:, :, 0] + 1j * mo[:, :, 1], bg


def test_tests_run_config(scene_json):
    """The reference's own end-to-end case (src/main_test.cpp:1884-2028): 20^3, 2 monitors, save_span 1.
    The reference only pins structure; it records Ex, which is identically zero for its Ey source."""
    series, ref, bg = _bound_geom_pair("tests_run", scene_json, "f64")
    assert len(bg.get_field_times()) == 2 == len(bg.get_monitor_locs())
    assert bg.n_t_pts == 234 and series.shape == (234, 2)          # 234 steps
    tb = bg.time_bounds()
    assert tb[0] == 0.0 and tb[0] < tb[2] < tb[1]
    # Ex is symmetry-suppressed (only edge/PML effects of the finite plane source feed it): tiny but not 0
    assert np.abs(ref).max() < 1e-4
    assert rel_l2(series, ref) < 1e-9


def test_cw_scene_through_bound_geom(scene_json):
    """CW_source end to end: own .geom reader -> bound_geom -> engine, against the oracle driven the same way.
    (scene: scenes/tests/cw_slab.geom; run length = CW end time + post_source_t, disp.cpp:625)"""
    import os
    from sim_juncs_b200.settings import ParseSettings
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    series, ref, bg = _bound_geom_pair("tests_cw_slab", scene_json, "f64")
    assert bg.get_sources()[0].type == "continuous" and series.shape[1] == 3
    assert np.abs(ref).max() > 1e-2 and rel_l2(series, ref) < 1e-11
    st = ParseSettings()
    conf = os.path.join(root, "scenes", "tests", "cw_slab.conf")
    st.parse_args(["--conf-file", conf])
    cwd = os.getcwd()
    os.chdir(root)                      # geom_fname in the conf is relative to the repo root
    try:
        st.parse_conf_file(conf)
        st.correct_defaults()
        bg2 = BoundGeom(st)             # scene read from st.geom_fname by sim_juncs_b200/cgs.py
        bg2.run()
    finally:
        os.chdir(cwd)
    assert np.array_equal(np.array(bg2.get_field_times()), np.array(bg.get_field_times()))


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_run_slabs_config(prec, scene_json):
    """C1': the intended geometry of tests/run.geom (two eps = 3.5 slabs), Ey source; the Ey series is
    compared because the reference's Ex monitors are identically zero here."""
    name = "tests_run_slabs"
    st = settings_from_doc(scene_json(name))
    sc = Scene.load(scene_json(name))
    bg = BoundGeom(st, scene_json(name), precision=prec, n_sets=2)          # GPU rasterization
    masks = [bg.sim.region_masks(c) for c in range(3)]
    o, n_t = oracle_bound_geom(sc, st, masks, nsets=2, mon_comp=1)
    s2 = Sim((20, 20, 20), st.resolution, pml=st.pml_thickness, n_sets=2, precision=prec)
    amb, reps, rpoles = region_tables(sc, st)
    s2.set_materials(materials_from_regions(amb, reps, rpoles), masks)
    info, (p1, p2) = sc.sources[0], sc.source_boxes[0]
    cba = LIGHT_SPEED * st.um_scale
    s2.add_gaussian_source(info.component, p1, p2, info.amplitude, 1 / (info.wavelen * st.um_scale), info.width * cba,
                           info.phase, info.start_time * cba, info.end_time * cba, True)
    s2.add_monitors(np.array(sc.monitor_locs), comp=1)
    s2.run(n_t, 1)
    o.run(n_t, 1)
    mo, mg = o.monitors(), s2.monitors()
    assert np.abs(mo).max() > 0.05
    assert rel_l2(mg, mo) < TOL[prec]


_GRAPHENE_ORACLE = {}


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_au_graphene_box_short(prec, scene_json):
    """configs[1] at its production grid (181^3, Drude Au + graphene sheet + Lorentz SiO2, complex fields):
    the first 560 steps -- the source switches on at step 482 -- against the oracle, whole E and H fields."""
    name, steps = "Au_graphene_box", 560
    st = settings_from_doc(scene_json(name))
    sc = Scene.load(scene_json(name))
    bg = BoundGeom(st, scene_json(name), precision=prec, n_sets=2)
    if "o" not in _GRAPHENE_ORACLE:
        masks = [bg.sim.region_masks(c) for c in range(3)]
        o, n_t = oracle_bound_geom(sc, st, masks, nsets=2)
        o.run(steps, st.save_span)
        _GRAPHENE_ORACLE["o"] = o
    o = _GRAPHENE_ORACLE["o"]
    bg.sim.run(steps, st.save_span)
    worst = 0.0
    for kind, off in (("E", 0), ("H", 3)):
        for q in range(2):
            scale = max(np.linalg.norm(o.field(kind, c, q)) for c in range(3))
            assert scale > 0
            for c in range(3):
                worst = max(worst, np.linalg.norm(bg.sim.field(off + c, q) - o.field(kind, c, q)) / scale)
    assert worst < TOL[prec], worst
    mo, mg = o.monitors(), bg.sim.monitors()
    assert mo.shape == mg.shape == (28, 50, 2)


@pytest.mark.parametrize("name", ["Au_SiO2_bowtie", "Au_SiO2_box", "quartz_box"])
def test_junction_scenes_reduced_grid(name, scene_json):
    """configs[2], [3], [4] (bowtie with the gold tips, box, quartz) on a 91^3 grid with the pulse moved to the start
    of the run, so that within 420 steps the wave has crossed the junction and entered the PML on every side: whole E
    and H fields and the monitor series against the oracle."""
    st = settings_from_doc(scene_json(name))
    st.grid_num = 91
    st.resolution = 91 / (2 * (st.len / 2 + st.pml_thickness))
    n = st.grid_cells()
    sc = early_pulse(Scene.load(scene_json(name)), 0.3)
    bg = BoundGeom(st, sc, n_sets=2)
    masks = [bg.sim.region_masks(c) for c in range(3)]
    o, _ = oracle_bound_geom(sc, st, masks, nsets=2)
    steps = 420
    o.run(steps, 10)
    bg.sim.run(steps, 10)
    worst = 0.0
    for kind, off in (("E", 0), ("H", 3)):
        for q in range(2):
            scale = max(np.linalg.norm(o.field(kind, c, q)) for c in range(3))
            assert scale > 1e-3, (kind, q, scale)
            for c in range(3):
                worst = max(worst, np.linalg.norm(bg.sim.field(off + c, q) - o.field(kind, c, q)) / scale)
    assert worst < TOL["f64"], worst
    mo, mg = o.monitors(), bg.sim.monitors()
    assert mo.shape == mg.shape and np.abs(mo).max() > 1e-4
    assert rel_l2(mg, mo) < TOL["f64"]
    assert n == 91


# ---------------------------------------------------------------- size-independent properties at full size
def test_slab_decomposition_is_bitwise(scene_json):
    """Two stacked z-slabs on one GPU with explicit halo exchange == the single-slab run, bit for bit
    (fields and monitors): the N-GPU path cannot change results."""
    name = "Au_graphene_box"
    st = settings_from_doc(scene_json(name))
    mk = lambda: early_pulse(Scene.load(scene_json(name)))      # the pulse must cross the cut within the run
    whole = BoundGeom(st, mk(), n_sets=1)
    n = st.grid_cells()
    cut = 49
    lo = BoundGeom(st, mk(), n_sets=1, kz=(0, cut))
    up = BoundGeom(st, mk(), n_sets=1, kz=(cut, n + 1))
    steps = 200
    whole.sim.run(steps, 5)
    for i in range(steps):
        if i % 5 == 0:
            lo.sim.sample()
            up.sim.sample()
        for s_ in (lo.sim, up.sim):
            s_.h_pass(0, n + 1)
        Sim.halo_exchange(lo.sim, up.sim, 0)
        for s_ in (lo.sim, up.sim):
            s_.e_pass(0, n + 1)
        Sim.halo_exchange(lo.sim, up.sim, 1)
        lo.sim.tick()
        up.sim.tick()
    lo.sim.sync()
    up.sim.sync()
    for comp in range(6):
        w = whole.sim.field(comp)
        assert np.array_equal(w[:cut], lo.sim.field(comp)) and np.array_equal(w[cut:], up.sim.field(comp)), comp
        if comp in (0, 4):                                        # Ex / Hy of the Ex-polarised pulse, both sides of the cut
            assert np.abs(w[cut - 2:cut]).max() > 1e-6 and np.abs(w[cut:cut + 20]).max() > 1e-6
    m = lo.sim.monitors() + up.sim.monitors()          # each rank fills the monitors it owns, 0 elsewhere
    assert np.array_equal(m, whole.sim.monitors())


def test_graph_replay_equals_direct_launches(scene_json):
    """sj_run replays one captured CUDA graph per step; stepping the same scene pass by pass through sj_pass /
    sj_tick launches the kernels directly.  Fields and monitors must agree bit for bit, across a drive-table
    reallocation (which re-captures the graph) as well."""
    name = "Au_graphene_box"
    st = settings_from_doc(scene_json(name))
    n = st.grid_cells()
    a = BoundGeom(st, early_pulse(Scene.load(scene_json(name))), n_sets=2)
    b = BoundGeom(st, early_pulse(Scene.load(scene_json(name))), n_sets=2)
    steps = 120
    a.sim.run(20, 5)
    a.sim.run(steps - 20, 5)                        # second call: the drive table grows, the graph is rebuilt
    for i in range(steps):
        if i % 5 == 0:
            b.sim.sample()
        b.sim.h_pass(0, n + 1)
        b.sim.e_pass(0, n + 1)
        b.sim.tick()
    b.sim.sync()
    assert a.sim.steps_done() == b.sim.steps_done() == steps
    for q in range(2):
        for comp in range(6):
            assert np.array_equal(a.sim.field(comp, q), b.sim.field(comp, q)), (comp, q)
    assert np.array_equal(a.sim.monitors(), b.sim.monitors())
    assert np.abs(a.sim.field(0, 0)).max() > 1e-6 and np.abs(a.sim.field(0, 1)).max() > 1e-6


def test_linearity_in_source_amplitude_at_full_size(scene_json):
    """The path is linear in the drive: at the full 181^3 bench size, a source 3x as strong gives monitor series
    3x as large (to round-off), and fields stay finite."""
    name = "Au_graphene_box"
    st = settings_from_doc(scene_json(name))
    a = BoundGeom(st, scene_json(name), n_sets=1)
    sc = Scene.load(scene_json(name))
    sc.sources[0].amplitude *= 3.0
    b = BoundGeom(st, sc, n_sets=1)
    for bg in (a, b):
        bg.sim.run(1600, 20)
    ma, mb = a.sim.monitors(), b.sim.monitors()
    assert np.isfinite(mb).all() and np.abs(ma).max() > 1e-6
    assert rel_l2(mb, 3.0 * ma) < 1e-13


def test_quadrature_sets_are_phase_shifted_copies(scene_json):
    """Linearity: the Im field set is the Re set of the same run with the source phase
    shifted by +pi/2."""
    name = "Au_SiO2_box"
    st = settings_from_doc(scene_json(name))
    a = BoundGeom(st, scene_json(name), n_sets=2)
    sc = Scene.load(scene_json(name))
    sc.sources[0].phase += np.pi / 2
    b = BoundGeom(st, sc, n_sets=1)
    for bg in (a, b):
        bg.sim.run(1600, 20)
    ma, mb = a.sim.monitors(), b.sim.monitors()
    assert np.abs(ma[:, :, 1]).max() > 1e-6
    assert rel_l2(mb[:, :, 0], ma[:, :, 1]) < 1e-12


def test_full_run_decays_and_stays_finite(scene_json):
    """Au_SiO2_box production run (217^3, 3506 steps, 1600 monitors): finite and bounded."""
    name = "Au_SiO2_box"
    st = settings_from_doc(scene_json(name))
    bg = BoundGeom(st, scene_json(name), precision="f32", n_sets=2)
    bg.run()
    assert bg.n_t_pts == 3506 and len(bg.get_field_times()) == 1600
    s = np.abs(np.array(bg.get_field_times()))
    assert np.isfinite(s).all() and s.max() < 1000
    assert s[:, -1].max() < s.max()            # the pulse has peaked at every monitor before the run ends


def test_phase_batch_equals_individual_runs(scene_json):
    """BASELINE config 5 (CEP sweep batched as independent field sets): a batch over phases
    [0, pi/2, 1.0] equals the complex run (Re, Im) and a separate run with the phase added."""
    name = "quartz_box"
    st = settings_from_doc(scene_json(name))
    st.grid_num, st.resolution = 101, 101 / 20.0           # 101^3: same scene, test-sized
    steps = 900
    batch = BoundGeom(st, scene_json(name), phases=[0.0, np.pi / 2, 1.0])
    cplx = BoundGeom(st, scene_json(name), n_sets=2)
    sc = Scene.load(scene_json(name))
    sc.sources[0].phase += 1.0
    single = BoundGeom(st, sc, n_sets=1)
    for bg in (batch, cplx, single):
        bg.sim.run(steps, 10)
    mb, mc, ms = batch.sim.monitors(), cplx.sim.monitors(), single.sim.monitors()
    assert np.abs(mc[:, :, 0]).max() > 1e-4
    assert np.array_equal(mb[:, :, 0], mc[:, :, 0])
    assert rel_l2(mb[:, :, 1], mc[:, :, 1]) < 1e-12
    assert rel_l2(mb[:, :, 2], ms[:, :, 0]) < 1e-12


def test_save_field_samples_npz(tmp_path, scene_json):
    st = settings_from_doc(scene_json("tests_run_slabs"))
    bg = BoundGeom(st, scene_json("tests_run_slabs"), n_sets=2)
    bg.run()
    path = bg.save_field_times(str(tmp_path))
    assert path.endswith("field_samples.h5")
    from sim_juncs_b200 import hdf5
    z = np.load(path[:-3] + ".npz")
    with hdf5.File(path) as f:                                   # the read pattern of scripts/time_space.py:17-49
        clusters = [k for k in f.keys() if "cluster" in k]
        points = list(f[clusters[0]].keys())[1:]
        assert clusters == ["cluster_0", "cluster_1"] and points == ["point_0", "point_1"]
        assert len(f[clusters[0]][points[0]]["time"]) == 234 and len(f[clusters[0]][points[0]]["frequency"]) == 128
        assert np.array_equal(np.array(f[clusters[0]][points[1]]["time"]["Re"]), z["cluster_0/point_1/time"][:, 0])
        assert np.array_equal(np.array(f[clusters[0]][points[1]]["frequency"]["Im"]), z["cluster_0/point_1/frequency"][:, 1])
        assert f["info"]["time_bounds"][1] == z["info/time_bounds"][1] and f["info"]["n_time_points"][0] == 234
        assert f["info"]["sources"][0]["wavelen"] == z["info/sources"][0, 0]
        assert f["info"]["cgs_params"]["l_per_um"][0] == 2
        assert f[clusters[0]]["locations"][1][2] == z["cluster_0/locations"][1, 2] and len(f["cluster_1"]["locations"]) == 0
    assert z["info/n_clusters"][0] == 1 and z["info/n_time_points"][0] == 234
    assert z["cluster_0/locations"].shape == (2, 3) and z["cluster_1/locations"].shape == (0, 3)
    assert z["cluster_0/point_0/time"].shape == (234, 2) and z["cluster_0/point_0/frequency"].shape == (128, 2)
    assert z["info/cgs_params/l_per_um"][0] == 2 and z["info/cgs_params/tot_len"][0] == 4     # main_test.cpp:1968-1977
    assert z["info/sources"].shape == (1, 6)


def test_device_spectra_equal_reference_transform(scene_json):
    """sj_read_spectra (warp-reduced DFT of the device-resident monitor series) against the host restatement of the
    reference's fft() quirks (2^k truncation, 2 pi / N of the full length, FFT order); 131 saves -> 128 bins."""
    from sim_juncs_b200.output import field_samples_dict, reference_fft
    name = "Au_SiO2_box"
    st = settings_from_doc(scene_json(name))
    st.grid_num = 91
    st.resolution = 91 / (2 * (st.len / 2 + st.pml_thickness))
    st.save_span = 2
    bg = BoundGeom(st, early_pulse(Scene.load(scene_json(name)), 0.3), n_sets=2)
    bg.run()
    n_saves = len(bg.get_field_times()[0])
    assert 128 <= n_saves < 256
    spec = bg.sim.spectra(0, 1)
    assert spec.shape == (len(bg.get_monitor_locs()), 128)
    peak = max(np.abs(reference_fft(f)).max() for f in bg.get_field_times())
    assert peak > 1e-4
    for j in (0, 7, len(bg.get_monitor_locs()) - 1):
        want = reference_fft(bg.get_field_times()[j])
        assert np.abs(spec[j] - want).max() < 1e-12 * peak
    real_only = bg.sim.spectra(0, None)
    assert np.abs(real_only[7] - reference_fft(np.array(bg.get_field_times()[7]).real)).max() < 1e-12 * peak
    d = field_samples_dict(bg)                       # the save path takes the device spectra
    f = d["cluster_00/point_0007/frequency"]
    assert np.array_equal(f[:, 0] + 1j * f[:, 1], spec[7])


def test_raw_dumps_eps_and_ex_files(tmp_path, scene_json):
    """run(out_dir) writes eps-000000.00.h5 (disp.cpp:696) and, with dump_raw, one ex-<time>.h5 per save (disp.cpp:732-737)
    with meep's dataset names and axis order (x slowest)."""
    import glob
    from sim_juncs_b200 import hdf5
    from sim_juncs_b200.output import centred
    st = settings_from_doc(scene_json("tests_run_slabs"))
    st.dump_raw = 1
    st.save_span = 13
    bg = BoundGeom(st, scene_json("tests_run_slabs"), n_sets=2)
    bg.run(str(tmp_path))
    n = st.grid_cells()
    eps = hdf5.File(str(tmp_path / "eps-000000.00.h5"))["eps"].read()
    assert eps.shape == (n, n, n)
    table = [m[0] for m in bg.sim.material_table()]
    present = sorted({table[m] for c in range(3) for m in np.unique(bg.sim.region_masks(c))})
    assert len(present) >= 2                                                             # slabs of eps 3.5 in the ambient
    assert present[0] - 1e-12 <= eps.min() < present[-1] and abs(eps.max() - present[-1]) < 1e-12    # the gap is one pixel wide
    assert len(np.unique(np.round(eps, 9))) > 2                                          # interface pixels are blended
    files = sorted(glob.glob(str(tmp_path / "ex-*.h5")))
    assert len(files) == (bg.n_t_pts + st.save_span - 1) // st.save_span == len(bg.get_field_times()[0])
    first, last = hdf5.File(files[0]), hdf5.File(files[-1])
    assert first.keys() == ["ex.i", "ex.r"] and not first["ex.r"].read().any()
    lr = last["ex.r"].read()
    assert lr.shape == (n, n, n) and np.abs(lr).max() > 1e-6
    # the run continued for the steps after the last save, so compare against a second run stopped at that save
    ref = BoundGeom(st, scene_json("tests_run_slabs"), n_sets=2)
    ref.sim.run(st.save_span * (len(files) - 1), st.save_span)
    assert np.array_equal(lr, centred(ref.sim.field(0, 0), 0))
    assert np.array_equal(last["ex.i"].read(), centred(ref.sim.field(0, 1), 0))
