"""save_field_times (reference src/disp.cpp:758-923) without libhdf5.

`<out_dir>/field_samples.h5` is written by the package's own HDF5 encoder (hdf5.py): the same groups,
dataset names, shapes and compound types ({Re,Im}, {x,y,z}, source_info at its C++ struct offsets) as the
reference.  The same datasets also go to `<out_dir>/field_samples.npz` with the HDF5 paths as keys (compound
types as 2-, 3- and 6-column arrays) for environments without any HDF5 reader.  `frequency` reproduces the reference's
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
    # `frequency`: on the device (sj_read_spectra, one warp per monitor and bin) when the series are the simulation's own
    spectra = None
    sim = getattr(bg, "sim", None)
    if with_frequency and sim is not None and hasattr(sim, "spectra") and getattr(bg, "phases", None) is None and len(bg.field_times) == n_locs \
            and n_locs and all(len(f) == sim.L.sj_n_samples(sim.h) for f in bg.field_times):
        spectra = sim.spectra(0, 1 if bg.n_sets >= 2 else None)
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
            # sic: the dataspace of `time` holds n_t_pts / save_span samples (disp.cpp:771) although the run loop pushed
            # ceil(n_t_pts / save_span); `frequency` below is the transform of everything pushed (disp.cpp:800-806)
            kept = series[:bg.n_t_pts // bg.save_span]
            d[pn + "/time"] = np.stack([kept.real, kept.imag], axis=1)
            if with_frequency:
                f = spectra[i] if spectra is not None else reference_fft(series)
                d[pn + "/frequency"] = np.stack([f.real, f.imag], axis=1)
            i += 1
    return d


FIELD_TYPE = np.dtype([("Re", "<f8"), ("Im", "<f8")])                       # disp.cpp:763-769, struct complex
LOC_TYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8")])             # disp.cpp:777-780, struct sto_vec
# disp.cpp:788-794: HOFFSET into class source_info (disp.hpp:96-109): two 4-byte enums, then six doubles
SRC_TYPE = np.dtype({"names": ["wavelen", "width", "phase", "start_time", "end_time", "amplitude"],
                     "formats": ["<f8"] * 6, "offsets": [8, 16, 24, 32, 40, 48], "itemsize": 56})


def _compound(a, dt):
    a = np.asarray(a, dtype=np.float64).reshape(-1, len(dt.names))
    out = np.zeros(len(a), dtype=dt)
    for c, name in enumerate(dt.names):
        out[name] = a[:, c]
    return out


def write_field_samples_h5(d, path):
    """The dict of field_samples_dict() as an HDF5 file with the reference's types."""
    from .hdf5 import H5Writer
    w = H5Writer(path)
    w.create_group("info/cgs_params")
    for key, val in d.items():
        leaf = key.rsplit("/", 1)[-1]
        if not key.startswith(("cluster", "info/sources")):
            pass
        elif leaf in ("time", "frequency"):
            val = _compound(val, FIELD_TYPE)
        elif leaf == "locations":
            val = _compound(val, LOC_TYPE)
        elif key == "info/sources":
            val = _compound(val, SRC_TYPE)
        w.create_dataset(key, val)
    w.close()
    return path


def save_field_samples(bg, out_dir, with_frequency=True):
    """-> path of field_samples.h5 (field_samples.npz is written beside it)."""
    os.makedirs(out_dir, exist_ok=True)
    d = field_samples_dict(bg, with_frequency)
    np.savez(os.path.join(out_dir, "field_samples.npz"), **d)
    return write_field_samples_h5(d, os.path.join(out_dir, "field_samples.h5"))


class _PhaseView:
    """one phase of a batched CEP sweep, shaped like the BoundGeom of a single complex-field run with that phase"""

    def __init__(self, bg, series, phase):
        import copy
        self.__dict__.update({k: getattr(bg, k) for k in ("monitor_clusters", "monitor_locs", "n_t_pts", "save_span", "problem",
                                                          "ttot", "um_scale")})
        self.time_bounds = bg.time_bounds
        self.field_times = series
        self.n_sets, self.phases, self.sim = 2, None, None
        self.sources = []
        for src in bg.sources:
            c = copy.copy(src)
            c.phase = src.phase + phase
            self.sources.append(c)


def save_phase_batch(bg, out_dir, indices=None):
    """A BoundGeom run with phases = [phi_0 .. phi_{m-1}, phi_0 + pi/2 .. phi_{m-1} + pi/2] (every phase with its
    quadrature, which is the imaginary part of meep's complex field for that phase): writes <out_dir>/phase_<index>/
    field_samples.h5 per phase -- the file a single run with that source phase writes.  Returns the paths."""
    m = len(bg.phases) // 2
    if indices is None:
        indices = list(range(m))
    paths = []
    for j in range(m):
        series = [np.asarray(bg.field_times[j][i]).real + 1j * np.asarray(bg.field_times[j + m][i]).real
                  for i in range(len(bg.field_times[j]))]
        view = _PhaseView(bg, series, bg.phases[j])
        paths.append(save_field_samples(view, os.path.join(out_dir, "phase_%03d" % indices[j])))
    return paths


# ---- whole-grid dumps: eps-000000.00.h5 at the start of run(), ex-<time>.h5 per save when dump_raw -------------
def make_dec_str(t, n_digits_a, n_digits_b, dec_char="."):
    """argparse.h:390-431: zero-padded `aaa.bbb`; None where the reference returns -2 (too many integer digits)."""
    bef = int(t)
    t -= float(bef)
    digits = []
    while True:
        if len(digits) == n_digits_a:
            return None
        digits.append(chr(ord("0") + bef % 10))
        bef //= 10
        if not bef:
            break
    out = "0" * (n_digits_a - len(digits)) + "".join(reversed(digits)) + dec_char
    for _ in range(n_digits_b):
        t *= 10
        out += chr(ord("0") + int(t))
        t -= math.floor(t)
    return out


def field_dump_name(time, ttot, dt):
    """disp.cpp:709-713,733: "ex-" + make_dec_str(fields.time(), ceil(log10 ttot), ceil(-log10 dt) + 1) + ".h5"."""
    n_digits_a = int(math.ceil(math.log(ttot) / math.log(10)))
    n_digits_b = int(math.ceil(-math.log(float(dt)) / math.log(10.0))) + 1
    return "ex-" + (make_dec_str(time, n_digits_a, n_digits_b) or "") + ".h5"


def centred(a, comp):
    """An E-component Yee array [k][j][i] of (n+1)^3 points -> (nx, ny, nz) values at the pixel centres, x slowest
    (meep's output grid and axis order): the mean of the four Yee points of that component around each centre."""
    if comp == 0:
        c = 0.25 * (a[:-1, :-1, :-1] + a[1:, :-1, :-1] + a[:-1, 1:, :-1] + a[1:, 1:, :-1])
    elif comp == 1:
        c = 0.25 * (a[:-1, :-1, :-1] + a[1:, :-1, :-1] + a[:-1, :-1, 1:] + a[1:, :-1, 1:])
    else:
        c = 0.25 * (a[:-1, :-1, :-1] + a[:-1, 1:, :-1] + a[:-1, :-1, 1:] + a[:-1, 1:, 1:])
    return np.ascontiguousarray(c.transpose(2, 1, 0))


def write_eps_h5(bg, out_dir):
    """fields.output_hdf5(meep::Dielectric, total_volume) (disp.cpp:696): dataset "eps" = n_c / sum_c <chi1inv_c> at the
    pixel centres, chi1inv_c = 1 / eps_inf at the Yee points of E component c (meep fields::get_eps, recalled)."""
    from .hdf5 import H5Writer
    table = np.array([m[0] for m in bg.sim.material_table()])
    tr = 0.0
    for c in range(3):
        tr = tr + centred(1.0 / table[bg.sim.material_ids(c)], c)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "eps-000000.00.h5")
    with H5Writer(path) as w:
        w.create_dataset("eps", 3.0 / tr)
    return path


def write_ex_h5(bg, out_dir, time):
    """fields.output_hdf5(meep::Ex, vol.surroundings(), file) per save (disp.cpp:732-737): datasets "ex.r", "ex.i"."""
    from .hdf5 import H5Writer
    path = os.path.join(out_dir, field_dump_name(time, bg.ttot, bg.sim.dt))
    with H5Writer(path) as w:
        w.create_dataset("ex.r", centred(bg.sim.field(0, 0), 0))
        if bg.n_sets >= 2 and bg.phases is None:
            w.create_dataset("ex.i", centred(bg.sim.field(0, 1), 0))
    return path
