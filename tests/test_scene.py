"""bound_geom's view of parsed scenes, pinned by reference src/main_test.cpp:1747-1873 (fixtures are the
reference parser's output, scripts/make_golden.py)."""
import numpy as np
import pytest

from sim_juncs_b200.scene import Scene


def test_susceptibilities_two_drude(scene_json):
    # main_test.cpp:1764-1777
    sc = Scene.load(scene_json("tests_test"))
    poles = sc.regions[0].poles_raw
    assert sc.regions[0].sus_ercode == 0 and len(poles) == 2
    assert poles[0] == (1.0, 0.48, 68.5971845, False)
    assert poles[1] == (8.0, 0.816, float(452848600800781300), False)
    assert sc.regions[0].eps == 3.5


def test_sources(scene_json):
    # main_test.cpp:1782-1805
    sc = Scene.load(scene_json("tests_test"))
    assert len(sc.sources) == 2
    g = sc.sources[0]
    assert g.type == "gaussian" and g.component == 1 and g.wavelen == 1.333333 and g.amplitude == 7.0
    assert g.width == pytest.approx(3.0) and g.phase == 0.75
    assert g.start_time == pytest.approx(5.2) and g.end_time == pytest.approx(41.2)   # 5.2 + 2*6*3.0: cutoff quirk
    c = sc.sources[1]
    assert c.type == "continuous" and c.component == 5 and c.wavelen == 1.66 and c.amplitude == 8.0
    assert c.phase == 0.0 and c.start_time == 0.2 and c.end_time == 1.2 and c.width == 0.1


def test_monitor_spans(scene_json):
    # main_test.cpp:1837-1870: tests/span.geom -> 2 clusters, 4 + 16 locations
    sc = Scene.load(scene_json("tests_span"))
    assert len(sc.monitor_clusters) == 2 and sc.monitor_clusters == [4, 20]
    locs = np.array(sc.monitor_locs)
    assert locs.shape == (20, 3)
    assert np.allclose(locs[:4, 0], [1, 1.5, 2, 2.5])
    k = 4
    for i in range(4):
        for j in range(4):
            assert locs[k] == pytest.approx([1 + 0.5 * j, 1 + 0.5 * i, 1.0])
            k += 1


def test_run_monitors_and_vacuum(scene_json):
    # main_test.cpp:1905-1912; tests/run.geom parses to a root without children
    sc = Scene.load(scene_json("tests_run"))
    assert sc.monitor_locs == [[1.0, 1.0, 1.0], [2.0, 4.0, 1.1]]
    assert len(sc.regions) == 1
    root = sc.nodes[sc.regions[0].root]
    assert root.child0 == -1 and root.child1 == -1
    params = dict(sc.cgs_params)
    assert params["l_per_um"] == 2 and params["tot_len"] == 4      # main_test.cpp:1968-1977


def test_au_loses_lorentz_pole(scene_json):
    # the second Au pole is written without commas -> parse stops after the Drude pole
    sc = Scene.load(scene_json("Au_SiO2_box"))
    assert len(sc.regions) == 2
    sio2, au = sc.regions                 # roots are stored newest first
    assert sio2.poles_raw == [(9.67865314895427, 0.08065544290795199, 1.12, True)]
    assert au.sus_ercode == -2 and au.poles_raw == [(1e-10, 0.04274738474121455, 4.0314052191361974e21, False)]
    assert len(sc.monitor_clusters) == 40 and len(sc.monitor_locs) == 1600
    assert len(sc.sources) == 1 and sc.sources[0].component == 0


def test_graphene_scene(scene_json):
    sc = Scene.load(scene_json("Au_graphene_box"))
    assert [len(r.poles_raw) for r in sc.regions] == [1, 1, 2]     # SiO2, graphene, Au (newest first)
    assert sc.regions[1].make_2d and not sc.regions[0].make_2d
    assert len(sc.monitor_locs) == 50
