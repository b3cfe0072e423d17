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


def oracle_raster_counts(scene, settings, comp):
    """in_bound's inner sums with smooth_n > 0: [n_regions, nz+1, ny+1, nx+1] counts and 8 * smooth_n."""
    L = orc.lib()
    nodes = scene.node_array()
    for reg in scene.regions:
        if reg.make_2d:
            thick = THICK_SCALE / settings.resolution
            for i in range(3):
                nodes[reg.root].M[3 * i + 2] = nodes[reg.root].M[3 * i + 2] * thick
    nreg = len(scene.regions)
    roots = (C.c_int32 * max(nreg, 1))(*[r.root for r in scene.regions])
    pts = np.ascontiguousarray(orc.smooth_points(settings.smooth_n, settings.smooth_rad))
    n = settings.grid_cells()
    out = np.zeros((max(nreg, 1), n + 1, n + 1, n + 1), dtype=np.uint8)
    L.csg_raster_counts(C.cast(nodes, C.c_void_p), roots, nreg, n, n, n, settings.resolution, comp,
                        pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out[:nreg], len(pts)


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
    if settings.smooth_n > 0:       # masks are ignored: in_bound averages over the smoothing offsets
        counts = [oracle_raster_counts(scene, settings, c) for c in range(3)]
        o.set_regions_counts(amb, reps, rpoles, counts[0][1], [c[0] for c in counts])
    else:
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


# ---- the reference's own driver (main.cpp + disp.cpp, compiled in place over oracle/shim) -------------------
REF_SIM_GEOM = os.path.join(ROOT, "oracle", "_ref", "sim_geom_ref")


def have_ref_sim_geom():
    """oracle/_ref/sim_geom_ref exists, or can be built because /root/reference is mounted."""
    if not os.path.exists(REF_SIM_GEOM) and os.path.isdir("/root/reference/src"):
        import subprocess
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    return os.path.exists(REF_SIM_GEOM)


def run_ref_sim_geom(conf, out_dir, extra=()):
    """Run the reference binary from the repository root; returns (entries, blob, stdout).  entries: the recorded
    HDF5 objects in creation order (oracle/shim/H5Cpp.h); dumps of eps / sigma land in out_dir."""
    import json
    import subprocess
    os.makedirs(out_dir, exist_ok=True)
    env = dict(os.environ, SJ_SHIM_DUMP=out_dir)
    res = subprocess.run([REF_SIM_GEOM, "--conf-file", conf, "--out-dir", out_dir] + list(extra), cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1200)
    if res.returncode != 0:
        raise RuntimeError("sim_geom_ref failed (%d):\n%s" % (res.returncode, res.stdout.decode()[-2000:]))
    with open(os.path.join(out_dir, "field_samples.h5.manifest.json")) as fp:
        entries = json.load(fp)["entries"]
    with open(os.path.join(out_dir, "field_samples.h5.manifest.bin"), "rb") as fp:
        blob = fp.read()
    return entries, blob, res.stdout.decode()


def ref_dataset(entries, blob, path):
    """One recorded dataset as a numpy array: f64 / u64 scalars as they are, compounds as (n, n_members) float64."""
    for e in entries:
        if e["what"] == "dataset" and e["path"] == path:
            raw = blob[e["offset"]:e["offset"] + e["nbytes"]]
            t = e["type"]
            if t["kind"] == "f64":
                return np.frombuffer(raw, dtype="<f8").copy()
            if t["kind"] == "u64":
                return np.frombuffer(raw, dtype="<u8").copy()
            dt = np.dtype({"names": [m["name"] for m in t["members"]], "formats": ["<f8"] * len(t["members"]),
                           "offsets": [m["offset"] for m in t["members"]], "itemsize": t["size"]})
            rec = np.frombuffer(raw, dtype=dt)
            return np.stack([rec[m["name"]] for m in t["members"]], axis=1) if len(rec) else np.zeros((0, len(t["members"])))
    raise KeyError(path)


def ref_series(entries, blob):
    """[n_saves][n_monitors] complex from the recorded cluster_*/point_*/time datasets, in monitor order."""
    cols = []
    for e in entries:
        if e["what"] == "dataset" and e["path"].endswith("/time"):
            a = ref_dataset(entries, blob, e["path"])
            cols.append(a[:, 0] + 1j * a[:, 1])
    return np.stack(cols, axis=1) if cols else np.zeros((0, 0), dtype=complex)
