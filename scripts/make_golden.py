"""Generates the committed fixtures from the REFERENCE's own scene-language code.

Runs only in the build container (needs /root/reference): compiles oracle/_ref/scene_dump from the
reference sources in place (oracle/Makefile target `ref`) and writes
  scenes/json/<name>.json            flattened scenes (reference parser output) used by tests, smoke
                                     and bench on the GPU box, where /root/reference does not exist
  tests/golden/masks_<name>.npz      composite_object::in() over the whole Yee grid, per E component
  tests/golden/inside_points.npz     10 000 LCG points + the six known-answer points of
                                     reference src/main_test.cpp:1556-1561 through tests/test.geom
Usage: python scripts/make_golden.py [scene-name ...]
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DUMP = os.path.join(ROOT, "oracle", "_ref", "scene_dump")
sys.path.insert(0, ROOT)
from sim_juncs_b200.settings import ParseSettings  # noqa: E402

RUN_SH_OPTS = "width=0.05;thick=0.2;inf_thick=0;wavelen=0.76;n_cycles=0.5"   # reference scripts/run.sh:76-78

# name -> (geom path, conf path, CLI args as the reference is launched, overrides after parsing)
CONFIGS = {
    # reference src/main_test.cpp:1884-1899 (conf, then pml/len/resolution overridden)
    "tests_run": (REF + "/tests/run.geom", REF + "/tests/run.conf", [], dict(pml_thickness=1.0, len=2.0, resolution=5.0)),
    "tests_run_slabs": (ROOT + "/scenes/tests/run_slabs.geom", REF + "/tests/run.conf", [], dict(pml_thickness=1.0, len=2.0, resolution=5.0)),
    "tests_span": (REF + "/tests/span.geom", REF + "/tests/span.conf", [], {}),
    "tests_test": (REF + "/tests/test.geom", REF + "/tests/test.conf", [], {}),
    # scripts/run.sh:78 production launch
    "Au_SiO2_box": (REF + "/junctions/Au_SiO2_box/junc.geom", REF + "/junctions/Au_SiO2_box/params.conf",
                    ["--grid-res", "12.0", "--opts", RUN_SH_OPTS], {}),
    "Au_SiO2_bowtie": (REF + "/junctions/Au_SiO2_bowtie/junc.geom", REF + "/junctions/Au_SiO2_bowtie/params.conf",
                       ["--grid-res", "12.0", "--opts", RUN_SH_OPTS], {}),
    "Au_graphene_box": (ROOT + "/scenes/Au_graphene_box/junc.geom", ROOT + "/scenes/Au_graphene_box/params.conf", [], {}),
    "quartz_box": (ROOT + "/scenes/quartz_box/junc.geom", ROOT + "/scenes/quartz_box/params.conf", [], {}),
    "tests_cw_slab": (ROOT + "/scenes/tests/cw_slab.geom", ROOT + "/scenes/tests/cw_slab.conf", [], {}),
    # scene-language feature sweep for the own parser (tests/test_cgs_parser.py)
    "parser_features": (ROOT + "/scenes/tests/parser_features.geom", REF + "/tests/run.conf", [],
                        dict(pml_thickness=1.0, len=4.0, um_scale=2.0, resolution=4.0)),
}
MASK_CONFIGS = ["tests_run_slabs", "Au_SiO2_box", "Au_SiO2_bowtie", "Au_graphene_box"]


def settings_for(name):
    geom, conf, argv, over = CONFIGS[name]
    s = ParseSettings()
    s.parse_args(argv)
    s.parse_conf_file(conf)
    s.correct_defaults()
    for k, v in over.items():
        setattr(s, k, v)
    return s, geom


def dump_args(s, geom, mode):
    return [DUMP, mode, geom, repr(s.pml_thickness), repr(s.len), repr(s.um_scale), s.out_dir or "/tmp",
            s.user_opts or "-", repr(s.resolution)]


def ref_points(s, geom, pts):
    """composite_object::in() of every root at pts (N,3) through the compiled reference."""
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in.f64"), os.path.join(td, "out.u8")
        np.ascontiguousarray(pts, dtype=np.float64).tofile(fi)
        subprocess.check_output(dump_args(s, geom, "points") + [fi, fo], cwd=REF)
        return np.fromfile(fo, dtype=np.uint8)


def yee_coords(n, a, comp):
    """h = m * (0.5 * inva) for the half-pixel index m of E component comp (kind 0 in csg_oracle.c)."""
    inva = 1.0 / a
    half = 0.5 * inva
    ax = [np.array([(2 * i + (1 if comp == d else 0)) * half for i in range(n + 1)]) for d in range(3)]
    z, y, x = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    os.makedirs(os.path.join(ROOT, "scenes", "json"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[1:]                      # optional: regenerate just these scene dumps
    for name in (only or CONFIGS):
        s, geom = settings_for(name)
        out = subprocess.check_output(dump_args(s, geom, "dump"), cwd=REF)
        doc = json.loads(out)
        doc["settings"] = {k: getattr(s, k) for k in ("pml_thickness", "len", "um_scale", "resolution", "grid_num", "n_dims",
                                                       "ambient_eps", "smooth_n", "post_source_t", "save_span", "user_opts")}
        with open(os.path.join(ROOT, "scenes", "json", name + ".json"), "w") as fp:
            json.dump(doc, fp, separators=(",", ":"))
        print(name, "ercode", doc["ercode"], "roots", doc["n_roots"], "grid", s.grid_cells(), "a", s.resolution)
    if only:
        return
    for name in MASK_CONFIGS:
        s, geom = settings_for(name)
        n, a = s.grid_cells(), s.resolution
        masks = []
        for comp in range(3):
            m = ref_points(s, geom, yee_coords(n, a, comp)).reshape(n + 1, n + 1, n + 1)
            masks.append(m)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "masks_%s.npz" % name), n=n, a=a, ex=masks[0], ey=masks[1], ez=masks[2])
        print("masks", name, n, [int(m.astype(bool).sum()) for m in masks])
    # inside-test points through tests/test.geom: LCG of reference src/main_test.hpp:17-33 style (our own constants)
    s, geom = settings_for("tests_test")
    state = 31415926
    pts = []
    for _ in range(30000):
        state = (state * 1103515245 + 12345) % (1 << 31)
        pts.append(0.3 + 0.6 * state / float(1 << 31))
    pts = np.array(pts).reshape(-1, 3)
    kat = np.array([[.45, .45, .45], [.45, .41, .6], [.65, .45, .41], [.71, .51, .51], [.55, .45, .45], [.55, .41, .85]])
    allp = np.concatenate([kat, pts])
    m = ref_points(s, geom, allp)
    print("KAT inside", m[:6].tolist(), "(reference main_test.cpp:1556-1561 expects 1,1,1,0,0,0)")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "inside_points.npz"), pts=allp, inside=m)


if __name__ == "__main__":
    main()
