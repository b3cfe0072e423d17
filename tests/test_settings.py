"""Config / CLI semantics pinned by the reference's own tests (src/main_test.cpp:1645-1745)."""
import os

import pytest

from sim_juncs_b200.settings import ParseSettings

CONF = """[simulation]
dimensions = 3
pml_thickness = 2.0
length = 16.0
um_scale = 1.0
resolution = 1.5
courant = 0.3
smooth_n = 1
smooth_rad = 0.25
save_span = 171

[misc]
post_source_t = 1.0
out_dir =/test_dir

[physical]
ambient_eps = 1.0

[junction]
middle_w = 4.0
junc_max_z = 8.0
geom_fname = tests/test.geom

[monitors]
near_rad = 0.2
"""


@pytest.fixture()
def conf(tmp_path):
    p = tmp_path / "test.conf"
    p.write_text(CONF)
    return str(p)


def test_conf_only(conf):
    # main_test.cpp:1649-1670
    a = ParseSettings()
    a.parse_conf_file(conf)
    assert a.n_dims == 3 and a.pml_thickness == 2.0 and a.len == 16.0 and a.um_scale == 1.0
    assert a.resolution == 1.55          # 1.5 -> 31 grid points over 20 -> 1.55 (odd-grid rounding)
    assert a.courant == 0.3 and a.smooth_n == 1 and a.smooth_rad == 0.25
    assert a.post_source_t == 1.0 and a.save_span == 171 and a.ambient_eps == 1.0
    assert a.geom_fname == "tests/test.geom" and a.out_dir == "/test_dir"


def test_cli_overrides(conf):
    # main_test.cpp:1672-1713
    a = ParseSettings()
    rest = a.parse_args(["./test", "--conf-file", "blah.conf", "--geom-file", "blah.geom", "--out-dir", "/blah",
                         "--grid-res", "3.0", "--length", "9.0", "--eps1", "2.0", "--opts", 'a = 0.1; b_option = [1,"blah"]'])
    assert rest == ["./test"]
    assert a.conf_fname == "blah.conf"
    a.parse_conf_file(conf)
    assert a.geom_fname == "blah.geom" and a.out_dir == "/blah" and a.len == 9.0
    assert a.resolution == pytest.approx(3.1538, abs=1e-4)
    assert a.ambient_eps == 2.0 and a.n_dims == 3 and a.pml_thickness == 2.0 and a.um_scale == 1.0
    assert a.courant == 0.3 and a.smooth_n == 1 and a.smooth_rad == 0.25 and a.post_source_t == 1.0
    assert a.user_opts == 'a = 0.1; b_option = [1,"blah"]'


@pytest.mark.parametrize("res,cells", [(12.0, 217), (14.0, 253), (10.0, 181), (56.0, 1009)])
def test_production_grids(res, cells):
    # length 16 + 2 pml -> odd grid counts; meep vol3d rounds L*a + 0.5
    a = ParseSettings()
    a.pml_thickness, a.len, a.resolution = 1.0, 16.0, res
    a.correct_defaults()
    assert a.grid_num == cells and a.grid_cells() == cells
    assert a.resolution == cells / 18.0


def test_negative_resolution_sentinel():
    # tests/run.conf carries no resolution: the -2.0 sentinel propagates
    a = ParseSettings()
    a.pml_thickness, a.len = 1.0, 3.0
    a.correct_defaults()
    assert a.grid_num == -7 and a.resolution == -7 / 5.0
