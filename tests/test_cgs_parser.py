"""The own CGS reader (sim_juncs_b200/cgs.py) against the reference parser.

Two kinds of evidence:
  * the known-answer cases of the reference's own unit tests (src/main_test.cpp:347-560 builtins,
    599-807 value parsing, 936-1043 operations, 1264-1347 context parsing, 1349-1446 file parsing),
    restated on our interpreter;
  * whole-file comparison, float for float, with the documents the compiled reference parser produced
    (scenes/json/*.json, written by scripts/make_golden.py).  Scene files kept in this repo are checked
    everywhere; the reference's own files only where /root/reference is mounted.
"""
import json
import math
import os

import pytest

from sim_juncs_b200 import cgs
from sim_juncs_b200.scene import Scene
from sim_juncs_b200.settings import ParseSettings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def pv(text, ctx=None):
    return (ctx or cgs.Context()).parse_value(text)


def err(text):
    with pytest.raises(cgs.CgsError) as e:
        pv(text)
    return e.value.code


# ---------------------------------------------------------------- builtins (main_test.cpp:347-560)
def test_range():
    assert pv("range(4)") == [0.0, 1.0, 2.0, 3.0]
    assert pv("range(1,4)") == [1.0, 2.0, 3.0]
    assert pv("range(1,4,0.5)") == [0.5 * i + 1 for i in range(6)]
    assert err("range()") == cgs.E_LACK_TOKENS
    assert err('range("1")') == cgs.E_BAD_TYPE
    assert err('range(0.5,"1")') == cgs.E_BAD_TYPE
    assert err('range(0.5,1,"2")') == cgs.E_BAD_TYPE


def test_linspace():
    assert pv("linspace(1,2,5)") == [1.0 + 0.25 * i for i in range(5)]
    assert pv("linspace(2,1,5)") == [2.0 - 0.25 * i for i in range(5)]
    assert err("linspace(2,1)") == cgs.E_LACK_TOKENS
    assert err("linspace(2,1,1)") == cgs.E_BAD_VALUE
    for bad in ('linspace("2",1,1)', 'linspace(2,"1",1)', 'linspace(2,1,"1")'):
        assert err(bad) == cgs.E_BAD_TYPE


def test_flatten():
    assert pv("flatten([])") == []
    assert pv("flatten([1,2,3])") == [1.0, 2.0, 3.0]
    assert pv("flatten([[1,2,3],[4,5],6])") == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]


def test_math_functions():
    assert pv("sin(3.1415926535/2)") == pytest.approx(1.0)
    assert pv("sin(3.1415926535/6)") == pytest.approx(0.5)
    assert pv("cos(3.1415926535/2)") == pytest.approx(0.0, abs=1e-9)
    assert pv("cos(3.1415926535/6)") == pytest.approx(pv("sqrt(3)/2"))
    assert pv("tan(3.141592653589793/4)") == pytest.approx(1.0)
    assert pv("tan(0)") == 0.0 and pv("exp(0)") == 1.0
    assert pv("exp(1)") == pytest.approx(2.718281828)
    for f in ("sin", "cos", "tan", "exp", "sqrt"):
        assert err(f + "()") == cgs.E_LACK_TOKENS
        assert err(f + '("a")') == cgs.E_BAD_TYPE


# ---------------------------------------------------------------- values (main_test.cpp:599-807)
def test_numbers():
    for text, want in (("1", 1.0), ("12", 12.0), (".25", 0.25), ("1.25", 1.25), (".25e10", 0.25e10), ("1.25e10", 1.25e10),
                       ("1.25e-10", 1.25e-10)):
        assert pv(text) == want


def test_strings():
    assert pv('"foo"') == "foo"
    assert pv('" foo bar "') == " foo bar "
    assert pv('"foo(bar)"') == "foo(bar)"
    assert pv('"foo\\"bar\\" "') == 'foo\\"bar\\" '


def test_lists():
    assert pv('["foo"]') == ["foo"]
    assert pv('["foo", 1]') == ["foo", 1.0]
    assert pv('[[1,2,3], ["4", "5", "6"]]') == [[1.0, 2.0, 3.0], ["4", "5", "6"]]
    assert pv("[[i*2 for i in range(2)], [i*2-1 for i in range(1,3)], [x for x in range(1,3,0.5)]]") == \
        [[0.0, 2.0], [1.0, 3.0], [1.0, 1.5, 2.0, 2.5]]
    assert pv("[[x*y for x in range(1,6)] for y in range(5)]") == [[float((x + 1) * y) for x in range(5)] for y in range(5)]


def test_vectors():
    v = pv("vec(1.2, 3.4,56.7)")
    assert isinstance(v, cgs.Vec3) and v.el == [1.2, 3.4, 56.7]


# ---------------------------------------------------------------- operations (main_test.cpp:936-1043)
def test_arithmetic_and_its_grouping():
    assert pv("1+1.1") == 2.1 and pv("2-1.25") == 0.75 and pv("2*1.1") == 2.2 and pv("2.2/2") == 1.1
    assert pv("1+3/2") == 2.5 and pv("(1+3)/2") == 2.0
    assert pv("2*9/4*3") == 1.5            # right-nested: 2*(9/(4*3)); the reference test expects exactly this
    assert pv("10-4-3") == 9.0             # and so 10-(4-3)
    assert pv("2^3^2") == 512.0


def test_comparisons():
    assert pv("2 == 2") == 1.0 and pv("1 == 2") == 0.0
    assert pv("[2, 3] == [2, 3]") == 1.0 and pv("[2, 3, 4] == [2, 3]") == 0.0 and pv("[2, 3, 4] == [2, 3, 5]") == 0.0
    assert pv('"apple" == "apple"') == 1.0 and pv('"apple" == "banana"') == 0.0


def test_string_concatenation():
    assert pv('"foo"+"bar"') == "foobar"


def test_ternary():
    assert pv("(1 == 2) ? 100 : 200") == 200.0 and pv("(2 == 2) ? 100 : 200") == 100.0
    assert pv("(1 > 2) ? 100 : 200") == 200.0 and pv("(1 < 2) ? 100 : 200") == 100.0


def test_undefined_names_are_zero_not_errors():
    assert pv("nothing_here + 2") == 2.0
    c = cgs.Context()
    c.emplace("half", 3.0)
    assert c.parse_value("-half + 1") == 1.0      # a leading sign is part of the (unknown) name "-half"
    assert c.parse_value("0-half + 1") == -4.0    # 0-(half+1): splits at the FIRST + or -


# ---------------------------------------------------------------- contexts (main_test.cpp:1264-1347)
def test_context_without_nesting():
    c = cgs.Context()
    c.read_from_lines(["a = 1", '"b"', 'c = ["d", "e"]'])
    assert c.lookup("a") == 1.0 and c.lookup("c") == ["d", "e"]
    assert c.fields[-2] == [None, "b"]


def test_context_with_nesting():
    c = cgs.Context()
    c.read_from_lines(['a = {name = "apple", values = [20, 11]}', "b = a.values[0]", "c = a.values[1] + a.values[0]+1"])
    a = c.lookup("a")
    assert isinstance(a, cgs.Inst) and a.lookup("name") == "apple" and a.lookup("values") == [20.0, 11.0]
    assert c.lookup("b") == 20.0 and c.lookup("c") == 32.0


def test_context_host_functions():
    def fun(ctx, args, names):
        if args[0] > 5:
            inst = cgs.Inst()
            inst.emplace("name", "hi")
            return inst
        return args[0]
    c = cgs.Context()
    c.emplace("test_fun", cgs.Func("test_fun", 1, fun))
    c.read_from_lines(["a = test_fun(1)", "b=test_fun(10)"])
    assert c.lookup("a") == 1.0 and c.lookup("b").lookup("name") == "hi"


# ---------------------------------------------------------------- files (main_test.cpp:1349-1446)
SCENE_TEXT = """Gaussian_source("Ex", 1.33, 1.0, 2.0, 0.0, 0.0, cutoff=.125, start_time=-1.25, Box([0,0,1], [4,4,1]))//remark glued on

CW_source("Hz", 0.625, 0.25, 0.75, end_time=2.25, slowness=12, Box([0,0,0], [1/2, 1, 1]))
Composite(eps = 3.5, alpha=1, color=10, [
    Box([1/2,0,0], [1, 1, 1]),
    Box([1/2,0,2], [2, 2, 3])
])
//a line holding nothing but a remark
/*and one that isn't over
until the next line*/
snapshot("/tmp/run_alpha.pgm", [1.2, 1.2, 1.2], look=[-1,-1,-1], scale=1, up=[0,0,-1]);
"""


def test_context_file_parsing():
    lines = cgs.split_lines(SCENE_TEXT)
    assert len(lines) == 10
    c = cgs.Context()
    cgs.setup_geometry_context(c)
    n0 = len(c.fields)
    c.read_from_lines(lines)
    assert len(c.fields) == n0 + 4
    g, cw, comp, snap = [f[1] for f in c.fields[n0:]]
    assert g.type == "Gaussian_source" and len(g.fields) == 9
    assert [g.lookup(k) for k in ("component", "wavelength", "amplitude", "width", "phase", "cutoff", "start_time")] == \
        [0.0, 1.33, 1.0, 2.0, 0.0, 0.125, -1.25]
    assert g.fields[-1][1].type == "Box"
    assert cw.type == "CW_source" and len(cw.fields) == 8
    assert [cw.lookup(k) for k in ("component", "wavelength", "amplitude", "start_time", "end_time", "slowness")] == \
        [5.0, 0.625, 0.25, 0.75, 2.25, 12.0]
    assert cw.fields[-1][1].type == "Box"
    assert comp.type == "Composite" and len(comp.fields) == 5
    assert [comp.lookup(k) for k in ("eps", "alpha", "color")] == [3.5, 1.0, 10.0]
    geom = comp.fields[-1][1]
    assert len(geom) == 2 and all(x.type == "Box" for x in geom)
    assert snap.type == "snapshot" and len(snap.fields) == 9
    assert snap.lookup("fname") == "/tmp/run_alpha.pgm" and snap.lookup("scale") == 1.0
    assert snap.lookup("res") == 255.0 and snap.lookup("n_samples") == 50000.0 and snap.lookup("step") == 0.01
    assert all(isinstance(snap.lookup(k), cgs.Vec3) for k in ("cam_v", "look_v", "up_v"))


def test_text_helpers():
    assert cgs.csv_to_list("a, b ,c") == ["a", "b", "c"]
    assert cgs.csv_to_list("a,[b,c],f(d,e),") == ["a", "[b,c]", "f(d,e)", ""]      # a trailing comma adds an element
    assert cgs.csv_to_list("") == []
    assert cgs.csv_to_list('"x y", z') == ['"x y"', "z"]
    assert cgs.strchr_block("f(a=1), b=2", "=") == 9
    assert cgs.token_block("[a for b in c] for d in e", "for") == 15
    assert cgs.token_block("formal", "for") == -1


# ---------------------------------------------------------------- whole files against the reference parser
def _settings_from_fixture(doc):
    st = ParseSettings()
    st.parse_args([])
    st.correct_defaults()
    for k in ("pml_thickness", "len", "um_scale", "resolution", "user_opts"):
        setattr(st, k, doc["settings"][k])
    return st


def _diff(a, b, path, out):
    if isinstance(a, dict) and isinstance(b, dict):
        for k in sorted(set(a) | set(b)):
            if k == "child_types":          # enum bookkeeping of the reference's node class, not scene content
                continue
            if k not in a or k not in b:
                out.append((path + "/" + k, "missing"))
            else:
                _diff(a[k], b[k], path + "/" + k, out)
    elif isinstance(a, list) and isinstance(b, list):
        if len(a) != len(b):
            out.append((path, "length", len(a), len(b)))
        else:
            for i, (x, y) in enumerate(zip(a, b)):
                _diff(x, y, "%s[%d]" % (path, i), out)
    elif isinstance(a, (int, float)) and isinstance(b, (int, float)) and not isinstance(a, bool):
        if float(a) != float(b):            # exact: the parser's arithmetic decides voxel boundaries
            out.append((path, a, b))
    elif a != b:
        out.append((path, a, b))


def _unsigned_zero(v):
    if isinstance(v, dict):
        return {k: _unsigned_zero(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unsigned_zero(x) for x in v]
    return 0.0 if (isinstance(v, float) and v == 0.0) else v


IN_REPO = {"tests_run_slabs": "scenes/tests/run_slabs.geom", "Au_graphene_box": "scenes/Au_graphene_box/junc.geom",
           "quartz_box": "scenes/quartz_box/junc.geom", "parser_features": "scenes/tests/parser_features.geom",
           "tests_cw_slab": "scenes/tests/cw_slab.geom"}
IN_REF = {"tests_run": "tests/run.geom", "tests_span": "tests/span.geom", "tests_test": "tests/test.geom",
          "Au_SiO2_box": "junctions/Au_SiO2_box/junc.geom", "Au_SiO2_bowtie": "junctions/Au_SiO2_bowtie/junc.geom"}


def _check_against_fixture(name, geom):
    with open(os.path.join(ROOT, "scenes", "json", name + ".json")) as fp:
        ref = json.load(fp)
    mine = cgs.parse_geom(geom, _settings_from_fixture(ref))
    out = []
    for k in ("ercode", "n_roots", "roots", "context"):
        _diff(mine[k], ref[k], k, out)
    assert out == [], out[:10]
    # and the two documents drive the host identically (the fixture's "%.17g" text keeps no sign on zero)
    a, b = Scene(_unsigned_zero(mine)), Scene(ref)
    assert len(a.nodes) == len(b.nodes) and len(a.regions) == len(b.regions)
    assert bytes(a.node_array()) == bytes(b.node_array())
    assert a.monitor_locs == b.monitor_locs and a.monitor_clusters == b.monitor_clusters and a.cgs_params == b.cgs_params
    assert [(s.type, s.component, s.wavelen, s.width, s.phase, s.amplitude, s.start_time, s.end_time) for s in a.sources] == \
        [(s.type, s.component, s.wavelen, s.width, s.phase, s.amplitude, s.start_time, s.end_time) for s in b.sources]
    return mine


@pytest.mark.parametrize("name", sorted(IN_REPO))
def test_own_parser_equals_reference_parser_on_repo_scenes(name):
    _check_against_fixture(name, os.path.join(ROOT, IN_REPO[name]))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's own .geom files exist only in the build container")
@pytest.mark.parametrize("name", sorted(IN_REF))
def test_own_parser_equals_reference_parser_on_reference_scenes(name):
    _check_against_fixture(name, os.path.join(REF, IN_REF[name]))


def test_feature_sweep_contents():
    """spot values of scenes/tests/parser_features.geom, so the fixture comparison above is not vacuous."""
    doc = _check_against_fixture("parser_features", os.path.join(ROOT, IN_REPO["parser_features"]))
    ctx = dict((k, v) for k, v in doc["context"]["__fields__"] if k)
    assert ctx["deep"] == 1.5 and ctx["chain"] == 9.0 and ctx["pw"] == 512.0 and ctx["label"] == "ab"
    assert ctx["tern_a"] == 11.0 and ctx["tern_b"] == 22.0 and ctx["last"] == 6.0 and ctx["from_bag"] == 8.0
    assert ctx["steps"] == [0.5, 1.0, 1.5, 2.0] and ctx["flat"] == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    assert doc["n_roots"] == 5
    assert [r["metadata"]["eps"] for r in doc["roots"]] == [5.0, 3.0, 1.5, 4.0, 2.25]      # newest first
    rot = doc["roots"][2]["tree"]["children"][0]            # Rotate(...) -> sub-composite whose primitives carry M
    m = rot["children"][0]["M"]
    assert m[0] == math.cos(math.pi / 5) and m[1] == -math.sin(math.pi / 5) and m[8] == 1.0
    comp = doc["roots"][4]["tree"]                           # Sphere, Complement[...], Cylinder, Plane, <empty>
    assert comp["children"][0]["type"] == "sphere"
    nxt = comp["children"][1]["children"]
    assert nxt[0]["type"] == "composite" and nxt[0]["invert"] == 0            # the complement's own content: not inverted
    assert nxt[1]["children"][0]["type"] == "cylinder" and nxt[1]["children"][0]["invert"] == 1   # its later siblings are


def test_failing_files_fail_the_same_way(tmp_path):
    p = tmp_path / "bad.geom"
    p.write_text("a = [1, 2\n")
    st = ParseSettings(); st.parse_args([]); st.correct_defaults()
    assert cgs.parse_geom(str(p), st)["ercode"] == cgs.E_BAD_SYNTAX
    p.write_text("b = 3\nc = b[0]\n")
    assert cgs.parse_geom(str(p), st)["ercode"] == cgs.E_BAD_TYPE
