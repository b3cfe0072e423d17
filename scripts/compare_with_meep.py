"""Diff a field_samples.h5 written by the REAL reference (sim_geom linked against libmeep v1.24) against one written by
this engine for the same launch line -- the check that turns "parity unpinned" into a number.

  python scripts/compare_with_meep.py <meep_run>/field_samples.h5 <our_run>/field_samples.h5 [--tol 1e-9]

Both files are read with the package's own HDF5 reader (sim_juncs_b200/hdf5.py; it reads libhdf5-written files, see
tests/test_hdf5.py).  Per monitor: relative L2 difference of the complex time series and of the `frequency` dataset;
then the info group (time bounds, number of samples, sources, monitor locations).  Exit status 0 if every series is
within --tol (north_star: 1e-9 in fp64 mode, 1e-4 in fp32 mode), 1 otherwise.

How to produce the two files on a machine that has meep v1.24 (README.md of the reference: cmake build of sim_geom):
    # reference
    ./sim_geom --conf-file <conf> --out-dir out_meep [--grid-res R] [--opts "..."]
    # this engine, same launch line (python host, or host/_ref/sim_geom / sim_geom_meep for the C++ hosts)
    python -m sim_juncs_b200 --conf-file <conf> --out-dir out_b200 [--grid-res R] [--opts "..."]
    python scripts/compare_with_meep.py out_meep/field_samples.h5 out_b200/field_samples.h5
Start with tests/run.conf (20^3, 234 steps) and scenes/tests/run_slabs.geom, then junctions/Au_SiO2_box at --grid-res 12.

If the difference is not round-off (<= 1e-9), the two modelling choices that were restated from recall are the suspects,
in this order:
 1. src_time::is_integrated.  The reference inherits meep's base-class default (believed true).  Re-run this engine with
    `--current-sources` (python host) / BoundGeom(integrated=False): the drive then enters as a current on D at
    t + dt/2 instead of a dipole moment in E = chi1inv (D - P - S).  The two differ by a half-step shift of the drive,
    visible at 1e-3..1e-2 of the series; whichever matches meep is the right one.
 2. Where chi1inv is sampled (meep structure::set_chi1inv evaluates eps at the centre of the pixel volume,
    ((h - q) + (h + q)) / 2, this engine at the Yee point h itself): identical except where a surface of the scene passes
    within one ulp of a Yee coordinate.  Compare the eps dumps first (eps-000000.00.h5 of both runs); a difference in a
    handful of voxels on a surface points here.
A residual at the 1e-6 level that grows with time instead points at the PML (profile exponent / R_asymptotic defaults),
one that is constant in time at the source weights of add_volume_source (edge weights of the plane source)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sim_juncs_b200 import hdf5  # noqa: E402


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a - b))


def series(f):
    """[(cluster, point, time complex[n], frequency complex[m] or None)] in file order"""
    out = []
    for cn in sorted(k for k in f.keys() if k.startswith("cluster_")):
        cl = f[cn]
        for pn in sorted(k for k in cl.keys() if k.startswith("point_")):
            t = cl[pn]["time"].read()
            fr = cl[pn]["frequency"].read() if "frequency" in cl[pn].keys() else None
            out.append((cn, pn, t["Re"] + 1j * t["Im"], None if fr is None else fr["Re"] + 1j * fr["Im"]))
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("reference_h5")
    ap.add_argument("ours_h5")
    ap.add_argument("--tol", type=float, default=1e-9)
    args = ap.parse_args()
    fr, fo = hdf5.File(args.reference_h5), hdf5.File(args.ours_h5)
    sr, so = series(fr), series(fo)
    ok = True
    if len(sr) != len(so):
        print("monitor count differs: reference %d, ours %d" % (len(sr), len(so)))
        ok = False
    worst_t = worst_f = 0.0
    for (cn, pn, tr, qr), (_, _, to, qo) in zip(sr, so):
        n = min(len(tr), len(to))
        if len(tr) != len(to):
            print("%s/%s: %d samples vs %d (comparing the first %d)" % (cn, pn, len(tr), len(to), n))
            ok = False
        dt = rel_l2(to[:n], tr[:n])
        df = rel_l2(qo[:min(len(qr), len(qo))], qr[:min(len(qr), len(qo))]) if qr is not None and qo is not None else float("nan")
        worst_t, worst_f = max(worst_t, dt), max(worst_f, df if df == df else 0.0)
        flag = "" if dt <= args.tol else "   <-- above tolerance"
        print("%s/%s  time rel-L2 %.3e  frequency rel-L2 %.3e  max|ref| %.3e%s" % (cn, pn, dt, df, float(np.abs(tr).max()), flag))
        ok &= dt <= args.tol
    for key in ("time_bounds", "n_time_points", "n_clusters"):
        a, b = fr["info"][key].read(), fo["info"][key].read()
        same = np.array_equal(a, b)
        print("info/%s: %s%s" % (key, "equal" if same else "DIFFERENT", "" if same else "  reference %s ours %s" % (a, b)))
        ok &= same
    print("worst: time %.3e, frequency %.3e, tolerance %.1e -> %s" % (worst_t, worst_f, args.tol, "PASS" if ok else "FAIL"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
