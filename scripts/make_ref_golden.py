"""Golden monitor series from the REFERENCE's own driver.

Runs only in the build container (needs /root/reference): oracle/_ref/sim_geom_ref is the reference's unmodified
main.cpp + disp.cpp + cgs*.cpp + data_utils.cpp compiled in place against oracle/shim/ (meep API slice over
oracle/fdtd_oracle.c, HDF5 recorder) -- see oracle/shim/meep.hpp.  For each case below the binary is launched with the
reference's own command line and what its save_field_times handed to HDF5 is stored as
  tests/golden/ref_<name>.npz   time (n_saves x n_monitors complex), frequency, locations, time_bounds, n_time_points,
                                sources, cgs_param names / values, the launch (conf, argv), sha256 of eps / sigma dumps
The -m gpu tests replay the same launch through BoundGeom on the CUDA engine and compare (tests/test_gpu_parity.py).
Usage: python scripts/make_ref_golden.py [name ...]
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

# name -> (conf, extra argv)
CASES = {
    "run_slabs": ("scenes/tests/run.conf", []),
    "cw_slab": ("scenes/tests/cw_slab.conf", []),
    "graphene_res2p5": ("scenes/tests/graphene_short.conf", ["--grid-res", "2.5"]),
    "graphene_short": ("scenes/tests/graphene_short.conf", []),
    "run_slabs_smooth1": ("scenes/tests/run_smooth.conf", []),          # stochastic boundary smoothing, dielectric
    "graphene_smooth2": ("scenes/tests/graphene_smooth.conf", []),      # smoothing of eps and of every pole's sigma
    "graphene_long": ("scenes/tests/graphene_long.conf", []),           # 5965 steps: late-time ringing, long after the pulse
    # the reference's own production launch (scripts/run.sh:78): its shipped Au_SiO2_box scene at --grid-res 12, 217^3 cells,
    # 3506 steps, 1600 monitors (about 20 minutes of CPU); the GPU test enters through scenes/json/Au_SiO2_box.json
    # quartz_box (BASELINE config 5, authored scene) on a coarse grid (not coarser: the 9.97 pole needs 2 pi f0 dt < 2, at
    # --grid-res 3 both engines blow up to 1e70 together); the GPU test runs it as a phase batch
    "quartz_res6": ("scenes/quartz_box/params.conf", ["--geom-file", "scenes/quartz_box/junc.geom", "--grid-res", "6.0"]),
    # the reference's shipped bowtie scene (BASELINE config 3) on a coarse grid, 73^3
    "bowtie_res4": ("/root/reference/junctions/Au_SiO2_bowtie/params.conf",
                    ["--geom-file", "/root/reference/junctions/Au_SiO2_bowtie/junc.geom", "--grid-res", "4.0",
                     "--opts", "width=0.05;thick=0.2;inf_thick=0;wavelen=0.76;n_cycles=0.5"]),
    "Au_SiO2_box_P": ("/root/reference/junctions/Au_SiO2_box/params.conf",
                      ["--geom-file", "/root/reference/junctions/Au_SiO2_box/junc.geom", "--grid-res", "12.0",
                       "--opts", "width=0.05;thick=0.2;inf_thick=0;wavelen=0.76;n_cycles=0.5"]),
}


MONITOR_STRIDE = {"graphene_long": 5, "Au_SiO2_box_P": 20, "bowtie_res4": 20}


def load_existing(out_dir):
    """(entries, blob) of a run that was made by hand: SJ_SHIM_DUMP=<dir> oracle/_ref/sim_geom_ref ... --out-dir <dir>"""
    import json
    with open(os.path.join(out_dir, "field_samples.h5.manifest.json")) as fp:
        entries = json.load(fp)["entries"]
    with open(os.path.join(out_dir, "field_samples.h5.manifest.bin"), "rb") as fp:
        return entries, fp.read()


def main(names, from_dir=None):
    assert helpers.have_ref_sim_geom(), "needs /root/reference"
    for name in names:
        conf, extra = CASES[name]
        with tempfile.TemporaryDirectory() as tmp:
            if from_dir:
                tmp = from_dir
                entries, blob = load_existing(from_dir)
            else:
                entries, blob, out = helpers.run_ref_sim_geom(conf, tmp, extra)
            d = {"conf": conf, "argv": np.array(extra, dtype=str)}
            d["time"] = helpers.ref_series(entries, blob)
            freq = [helpers.ref_dataset(entries, blob, e["path"]) for e in entries
                    if e["what"] == "dataset" and e["path"].endswith("/frequency")]
            d["frequency"] = np.stack([f[:, 0] + 1j * f[:, 1] for f in freq], axis=1)
            d["locations"] = np.concatenate([helpers.ref_dataset(entries, blob, e["path"]) for e in entries
                                             if e["what"] == "dataset" and e["path"].endswith("/locations")])
            for k in ("time_bounds", "n_time_points", "n_clusters", "sources"):
                d[k] = helpers.ref_dataset(entries, blob, "/info/" + k)
            cg = [e["path"] for e in entries if e["what"] == "dataset" and e["path"].startswith("/info/cgs_params/")]
            d["cgs_names"] = np.array([p.rsplit("/", 1)[1] for p in cg], dtype=str)
            d["cgs_values"] = np.array([helpers.ref_dataset(entries, blob, p)[0] for p in cg])
            for f in sorted(os.listdir(tmp)):
                if f.endswith(".f64"):
                    with open(os.path.join(tmp, f), "rb") as fp:
                        d["sha256_" + f[:-4]] = np.array(hashlib.sha256(fp.read()).hexdigest())
            # small grids: keep eps itself (the smooth_n > 0 fixture is compared value by value)
            eps = [np.fromfile(os.path.join(tmp, "eps_%s.f64" % c)) for c in "xyz"] if os.path.exists(os.path.join(tmp, "eps_x.f64")) else [np.zeros(10 ** 6)]
            if eps[0].size <= 30 ** 3:
                d["eps"] = np.stack(eps)
                sig = sorted(f for f in os.listdir(tmp) if f.startswith("sigma_"))
                if sig:
                    d["sigma"] = np.stack([np.fromfile(os.path.join(tmp, f)) for f in sig])
            if name in MONITOR_STRIDE:              # long runs: keep every n-th monitor's series, all locations
                d["time"] = d["time"][:, ::MONITOR_STRIDE[name]]
                d["frequency"] = d["frequency"][:, ::MONITOR_STRIDE[name]]
                d["monitor_stride"] = np.array(MONITOR_STRIDE[name])
            path = os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name)
            np.savez_compressed(path, **d)
            print(name, d["time"].shape, "max |Ex| %.3e" % np.abs(d["time"]).max(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--from-dir":       # --from-dir <dir> <name>: use the output of a run made by hand
        main(sys.argv[3:4], from_dir=sys.argv[2])
    else:
        main(sys.argv[1:] or [n for n in CASES if n != "Au_SiO2_box_P"])
