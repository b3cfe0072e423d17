"""Pins the CSG restatement (oracle/csg_oracle.c) against the reference:
  * the six known-answer inside tests of reference src/main_test.cpp:1556-1561,
  * the eps lookups of src/main_test.cpp:1635-1640 (3.5 inside / 1.0 outside),
  * 10 000 points and whole Yee grids evaluated by the COMPILED reference
    (oracle/_ref/scene_dump, fixtures written by scripts/make_golden.py) -- bit for bit."""
import numpy as np
import pytest

from helpers import oracle_points, oracle_raster, settings_from_doc
from sim_juncs_b200.materials import materials_from_regions
from sim_juncs_b200.scene import Scene

KAT = np.array([[.45, .45, .45], [.45, .41, .6], [.65, .45, .41], [.71, .51, .51], [.55, .45, .45], [.55, .41, .85]])


def test_inside_known_answers(scene_json):
    sc = Scene.load(scene_json("tests_test"))
    assert oracle_points(sc, KAT).tolist() == [1, 1, 1, 0, 0, 0]


def test_eps_lookup_known_answers(scene_json):
    # cgs_material_function(root) with one region: in_bound = def + (eps - def) * in  -> 3.5 / 1.0
    sc = Scene.load(scene_json("tests_test"))
    pts = np.array([[.45, .45, .45], [.45, .41, .6], [.65, .45, .45], [.71, .51, .51], [.55, .45, .45], [.55, .41, .85]])
    mats = materials_from_regions(1.0, [sc.regions[0].eps], [[]])
    eps = [mats[m][0] for m in oracle_points(sc, pts)]
    assert eps == [3.5, 3.5, 3.5, 1.0, 1.0, 1.0]


def test_points_vs_compiled_reference(scene_json, golden):
    g = np.load(golden + "/inside_points.npz")
    sc = Scene.load(scene_json("tests_test"))
    got = oracle_points(sc, g["pts"])
    assert np.array_equal(got, g["inside"])
    assert g["inside"][:6].tolist() == [1, 1, 1, 0, 0, 0]
    assert 50 < int(g["inside"].sum()) < len(g["inside"])     # the sample really straddles the object


@pytest.mark.parametrize("name", ["tests_run_slabs", "Au_SiO2_box", "Au_SiO2_bowtie", "Au_graphene_box"])
def test_grid_masks_vs_compiled_reference(name, scene_json, golden):
    g = np.load(golden + "/masks_%s.npz" % name)
    sc = Scene.load(scene_json(name))
    st = settings_from_doc(scene_json(name))
    assert st.grid_cells() == int(g["n"])
    for comp, key in enumerate(("ex", "ey", "ez")):
        got = oracle_raster(sc, st, comp)
        assert np.array_equal(got, g[key]), "component %s differs in %d points" % (key, int((got != g[key]).sum()))


def test_empty_scene_is_vacuum(scene_json):
    # tests/run.geom: root without children -> in() == invert == 0 everywhere (cgs.cpp:426-428)
    sc = Scene.load(scene_json("tests_run"))
    assert oracle_points(sc, np.random.default_rng(0).random((100, 3)) * 4).sum() == 0


# ---- stochastic boundary smoothing (smooth_n > 0, reference src/disp.cpp:56-112, 264-283) -------------------------
def test_smooth_points_equal_libstdcxx(tmp_path):
    """The restated seed_seq / mt19937 / generate_canonical / polar normal_distribution against the C++ standard
    library itself (tests/cpp/smooth_pts_main.cpp calls <random> in the reference's order), bit for bit."""
    import os
    import shutil
    import subprocess
    from helpers import ROOT, orc
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "smooth_pts")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "smooth_pts_main.cpp")])
    for n, rad in ((1, 0.05), (7, 0.15), (31, 1.0)):
        want = np.array([[float.fromhex(t) for t in line.split()] for line in
                         subprocess.check_output([exe, str(n), repr(rad)]).decode().splitlines()])
        got = orc.smooth_points(n, rad)
        assert got.shape == (8 * n, 3)
        # reflection order: x sign slowest, z sign fastest, (-,-,-) first and (+,+,+) last
        assert np.array_equal(got[7::8], want) and np.array_equal(got[0::8], -want)
        assert np.array_equal(got[1::8], want * [-1, -1, 1]) and np.array_equal(got[4::8], want * [1, -1, -1])
        assert np.array_equal(want[:, 0], want[:, 1])                  # sic: y uses cos(phi) like x (disp.cpp:92)


def test_smoothed_eps_equals_reference_in_bound(golden):
    """Oracle counts + in_bound's sum against the eps field the reference's own cgs_material_function returned
    (tests/golden/ref_run_slabs_smooth1.npz, from oracle/_ref/sim_geom_ref), every Yee point, exact."""
    import os
    from helpers import ROOT, oracle_bound_geom
    from sim_juncs_b200.settings import settings_from
    g = np.load(golden + "/ref_run_slabs_smooth1.npz")
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        st = settings_from(str(g["conf"]), [str(a) for a in g["argv"]])
        sc = Scene.from_geom(st.geom_fname, st)
        o, n_t_pts = oracle_bound_geom(sc, st, None)
    finally:
        os.chdir(cwd)
    assert st.smooth_n == 1
    n = st.grid_cells()
    levels = set()
    for c in range(3):
        eps = g["eps"][c].reshape(n + 1, n + 1, n + 1)
        assert np.array_equal(1.0 / eps, o.field("chi1inv", c))
        levels |= set(np.unique(eps).tolist())
    assert len(levels) > 2 and min(levels) == 1.0 and max(levels) == 3.5     # partial levels exist: 1 + 2.5 c / 9
