"""Shared test helpers: the oracle is the checker, the CUDA engine is the thing under test."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from sim_juncs_b200.scene import LIGHT_SPEED, THICK_SCALE, Scene  # noqa: E402
from sim_juncs_b200.settings import ParseSettings  # noqa: E402


def settings_from_doc(path):
    """Rebuild the ParseSettings a fixture was generated with (scripts/make_golden.py stores them)."""
    import json
    with open(path) as fp:
        st = json.load(fp)["settings"]
    s = ParseSettings()
    for k, v in st.items():
        setattr(s, k, v)
    s.out_dir = "/tmp"
    return s


def oracle_csg_nodes(scene):
    """Scene nodes -> oracle csg_node array (same binary layout as SjCsgNode)."""
    return scene.node_array()


def oracle_raster(scene, settings, comp, coord_kind=0):
    L = orc.lib()
    L.csg_raster_component.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                       C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    nodes = scene.node_array()
    # make_2d rescale (disp.cpp:521-524) applied to a copy of the root matrix
    for reg in scene.regions:
        if reg.make_2d:
            thick = THICK_SCALE / settings.resolution
            for i in range(3):
                nodes[reg.root].M[3 * i + 2] = nodes[reg.root].M[3 * i + 2] * thick
    roots = (C.c_int32 * max(len(scene.regions), 1))(*[r.root for r in scene.regions])
    n = settings.grid_cells()
    out = np.zeros((n + 1, n + 1, n + 1), dtype=np.uint8)
    L.csg_raster_component(C.cast(nodes, C.c_void_p), roots, len(scene.regions), n, n, n, settings.resolution, comp,
                           coord_kind, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def oracle_points(scene, pts):
    L = orc.lib()
    L.csg_eval_points.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_double), C.c_size_t,
                                  C.POINTER(C.c_uint8)]
    nodes = scene.node_array()
    roots = (C.c_int32 * max(len(scene.regions), 1))(*[r.root for r in scene.regions])
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.zeros(len(pts), dtype=np.uint8)
    L.csg_eval_points(C.cast(nodes, C.c_void_p), roots, len(scene.regions), pts.ctypes.data_as(C.POINTER(C.c_double)),
                      len(pts), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def region_tables(scene, settings):
    """(ambient, region_eps, region_poles in meep units) exactly as structure_from_settings does."""
    reps, rpoles = [], []
    for reg in scene.regions:
        thick = THICK_SCALE / settings.resolution if reg.make_2d else 1.0
        reps.append(reg.eps if reg.eps is not None else settings.ambient_eps)
        rpoles.append([(w0 / settings.um_scale, g / settings.um_scale, sg / thick, 0 if use_denom else 1)
                       for (w0, g, sg, use_denom) in reg.poles_raw])
    return settings.ambient_eps, reps, rpoles


def oracle_bound_geom(scene, settings, masks, nsets=2, integrated=True, mon_comp=0):
    """The oracle driven the way bound_geom drives meep (disp.cpp:557-645): returns (sim, n_t_pts)."""
    n = settings.grid_cells()
    o = orc.OracleSim((n, n, n), settings.resolution, pml=settings.pml_thickness, nsets=nsets)
    amb, reps, rpoles = region_tables(scene, settings)
    o.set_regions(amb, reps, rpoles, masks)
    c_by_a = LIGHT_SPEED * settings.um_scale
    ttot = 0.0
    for info, (p1, p2) in zip(scene.sources, scene.source_boxes):
        lo = [min(p1[d], p2[d]) for d in range(3)]
        hi = [max(p1[d], p2[d]) for d in range(3)]
        if info.type == "gaussian":
            o.add_gaussian_source(info.component, lo, hi, info.amplitude, 1 / (info.wavelen * settings.um_scale),
                                  info.width * c_by_a, info.phase, info.start_time * c_by_a, info.end_time * c_by_a, integrated)
        else:       # disp.cpp:615-619
            o.add_cw_source(info.component, lo, hi, info.amplitude, 1 / (info.wavelen * settings.um_scale),
                            info.width * c_by_a, info.start_time * c_by_a, info.end_time * c_by_a, 3.0, integrated)
        ttot = o.last_source_time() + settings.post_source_t * LIGHT_SPEED * settings.um_scale
    if scene.monitor_locs:
        o.add_monitors(np.array(scene.monitor_locs), mon_comp)
    n_t_pts = int((ttot + o.dt / 2) / o.dt)
    return o, n_t_pts


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / nb if nb > 0 else np.linalg.norm(a - b)


def early_pulse(scene, width_fs=0.08):
    """Move the first source's pulse to the start of the run (start 0, peak at 6 widths) so that short test runs carry
    a non-zero field: the shipped scenes start their pulse at 5 fs, hundreds of steps in."""
    src = scene.sources[0]
    src.start_time, src.width, src.end_time = 0.0, width_fs, 12.0 * width_fs
    return scene
