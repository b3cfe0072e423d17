"""save_field_times (reference src/disp.cpp:758-923) without libhdf5.

HDF5 is not available in this image, so the same datasets are written to
`<out_dir>/field_samples.npz` with the HDF5 paths as keys; compound types become 2-column arrays
({Re,Im}, {x,y,z}) and a 6-column array for /info/sources.  `frequency` reproduces the reference's
transform (src/data_utils.cpp:370-427): only the first 2^floor(log2 N) samples are used while the
phase step stays 2 pi / N of the FULL length, output in FFT order.
"""
import math
import os

import numpy as np


def reference_fft(series):
    """data_utils.cpp fft(): X[k] = sum_{n < T} x[n] exp(-i (2 pi / N) n k), T = 2^floor(log2 N), k in FFT order."""
    x = np.asarray(series, dtype=np.complex128)
    n_full = len(x)
    if n_full == 0:
        return np.zeros(0, dtype=np.complex128)
    log_2_n = int(math.log(n_full) / math.log(2))          # same libm expression as the reference
    t = 1 << log_2_n
    tau_by_n = 2 * math.pi / n_full
    half = t // 2
    ks = np.concatenate([np.arange(0, half), np.arange(-half, 0)]) if t > 1 else np.array([0])
    n = np.arange(t)
    out = np.exp(-1j * tau_by_n * np.outer(ks, n)) @ x[:t]
    return out


def point_name(i, n_locs):
    digits = int(math.log(n_locs) / math.log(10)) + 1
    return "point_" + str(i).zfill(digits)


def cluster_name(j, n_clusters):
    digits = int(math.log(n_clusters) / math.log(10)) + 1 if n_clusters > 0 else 1
    return "cluster_" + str(j).zfill(digits)


def field_samples_dict(bg, with_frequency=True):
    """Every dataset of field_samples.h5, keyed by its HDF5 path."""
    d = {}
    d["info/time_bounds"] = np.array(bg.time_bounds())
    d["info/n_clusters"] = np.array([len(bg.monitor_clusters)], dtype=np.uint64)
    d["info/n_time_points"] = np.array([bg.n_t_pts // bg.save_span], dtype=np.uint64)
    if bg.sources:
        d["info/sources"] = np.array([[s.wavelen, s.width, s.phase, s.start_time, s.end_time, s.amplitude] for s in bg.sources])
    for name, val in bg.problem.cgs_params:
        d["info/cgs_params/" + name] = np.array([val])
    locs = np.array(bg.monitor_locs, dtype=np.float64).reshape(-1, 3)
    n_locs = len(locs)
    n_cl = len(bg.monitor_clusters)
    i = off = 0
    for j in range(n_cl + 1):                 # sic: the reference writes one empty trailing cluster (disp.cpp:879)
        max_i = n_locs if j >= n_cl else bg.monitor_clusters[j]
        cn = cluster_name(j, n_cl)
        d[cn + "/locations"] = locs[off:max_i]
        off = max_i
        while i < max_i:
            series = np.asarray(bg.field_times[i])
            if len(series) < 2:
                break
            pn = cn + "/" + point_name(i, n_locs)
            d[pn + "/time"] = np.stack([series.real, series.imag], axis=1)
            if with_frequency:
                f = reference_fft(series)
                d[pn + "/frequency"] = np.stack([f.real, f.imag], axis=1)
            i += 1
    return d


def save_field_samples(bg, out_dir, with_frequency=True):
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "field_samples.npz")
    np.savez(path, **field_samples_dict(bg, with_frequency))
    return path
