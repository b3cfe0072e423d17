"""ctypes front end for the CPU oracle (oracle/fdtd_oracle.c, oracle/csg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by sim_juncs_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in ("fdtd_oracle.c", "csg_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        dp = C.POINTER(C.c_double)
        u8p = C.POINTER(C.c_uint8)
        ip = C.POINTER(C.c_int)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_regions.argtypes = [C.c_void_p, C.c_double, C.c_int, dp, ip, dp, u8p, u8p, u8p]
        L.orc_set_regions_counts.argtypes = [C.c_void_p, C.c_double, C.c_int, dp, ip, dp, C.c_uint, u8p, u8p, u8p]
        L.csg_smooth_points.argtypes = [C.c_int, C.c_double, dp]
        L.csg_raster_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_int, dp, C.c_int, u8p]
        L.orc_add_gaussian_source.argtypes = [C.c_void_p, C.c_int, dp, dp] + [C.c_double] * 7 + [C.c_int]
        L.orc_add_cw_source.argtypes = [C.c_void_p, C.c_int, dp, dp] + [C.c_double] * 7 + [C.c_int]
        L.orc_add_monitors.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_step.argtypes = [C.c_void_p]
        L.orc_n_saves.argtypes = [C.c_void_p]
        L.orc_read_monitors.argtypes = [C.c_void_p, dp]
        L.orc_field.restype = dp
        L.orc_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_pol.restype = dp
        L.orc_pol.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_dt.restype = C.c_double
        L.orc_dt.argtypes = [C.c_void_p]
        L.orc_src_last_time.restype = C.c_double
        L.orc_src_last_time.argtypes = [C.c_void_p]
        L.orc_src_dipole.argtypes = [C.c_void_p, C.c_int, C.c_double, dp]
        L.orc_pml_sig.restype = dp
        L.orc_pml_sig.argtypes = [C.c_void_p, C.c_int]
        L.orc_src_range.argtypes = [C.c_void_p, C.c_int, ip, ip]
        L.orc_src_weights.restype = dp
        L.orc_src_weights.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_src_amp.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def smooth_points(smooth_n, smooth_rad):
    """generate_smooth_pts (disp.cpp:56-112): (8 * smooth_n, 3) offsets, restated in csg_oracle.c."""
    out = np.zeros(24 * max(int(smooth_n), 0))
    if smooth_n > 0:
        lib().csg_smooth_points(int(smooth_n), float(smooth_rad), _dp(out))
    return out.reshape(-1, 3)


class OracleSim:
    """Thin object wrapper; array layout is [k][j][i] with shape (nz+1, ny+1, nx+1)."""

    KIND = {"E": 0, "H": 1, "D": 2, "B": 3, "W": 4, "chi1inv": 5}

    def __init__(self, n, a, pml=0.0, courant=0.5, R=1e-15, nsets=1):
        self.L = lib()
        self.n = tuple(int(x) for x in n)
        self.nsets = nsets
        self.shape = (self.n[2] + 1, self.n[1] + 1, self.n[0] + 1)
        self.h = self.L.orc_create(self.n[0], self.n[1], self.n[2], a, courant, pml, R, nsets)
        self.n_mon = 0
        self.n_src = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    @property
    def dt(self):
        return self.L.orc_dt(self.h)

    def set_regions(self, ambient_eps, region_eps, region_poles, masks):
        """region_poles: list (per region) of lists of (omega0, gamma, sigma, drude)."""
        nreg = len(region_eps)
        eps = np.ascontiguousarray(region_eps, dtype=np.float64)
        npoles = np.ascontiguousarray([len(p) for p in region_poles], dtype=np.int32)
        flat = np.ascontiguousarray([x for p in region_poles for pole in p for x in pole], dtype=np.float64)
        if flat.size == 0:
            flat = np.zeros(4)
        ms = [np.ascontiguousarray(m, dtype=np.uint8).reshape(self.shape) for m in masks]
        self._keep = (eps, npoles, flat, ms)
        u8p = C.POINTER(C.c_uint8)
        rc = self.L.orc_set_regions(
            self.h, float(ambient_eps), nreg, _dp(eps), npoles.ctypes.data_as(C.POINTER(C.c_int)), _dp(flat),
            ms[0].ctypes.data_as(u8p), ms[1].ctypes.data_as(u8p), ms[2].ctypes.data_as(u8p))
        if rc:
            raise RuntimeError("orc_set_regions failed: %d" % rc)

    def set_regions_counts(self, ambient_eps, region_eps, region_poles, smooth_total, counts):
        """counts: per E component an array [n_regions, nz+1, ny+1, nx+1] of in_bound's inner sums (smooth_n > 0)."""
        nreg = len(region_eps)
        eps = np.ascontiguousarray(region_eps, dtype=np.float64)
        npoles = np.ascontiguousarray([len(p) for p in region_poles], dtype=np.int32)
        flat = np.ascontiguousarray([x for p in region_poles for pole in p for x in pole], dtype=np.float64)
        if flat.size == 0:
            flat = np.zeros(4)
        cs = [np.ascontiguousarray(c, dtype=np.uint8).reshape((nreg,) + self.shape) for c in counts]
        u8p = C.POINTER(C.c_uint8)
        rc = self.L.orc_set_regions_counts(
            self.h, float(ambient_eps), nreg, _dp(eps), npoles.ctypes.data_as(C.POINTER(C.c_int)), _dp(flat),
            int(smooth_total), cs[0].ctypes.data_as(u8p), cs[1].ctypes.data_as(u8p), cs[2].ctypes.data_as(u8p))
        if rc:
            raise RuntimeError("orc_set_regions_counts failed: %d" % rc)

    def add_gaussian_source(self, comp, lo, hi, amp, freq, width, phase, t_start, t_end, integrated=True):
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        amp = complex(amp)
        rc = self.L.orc_add_gaussian_source(self.h, comp, _dp(lo), _dp(hi), amp.real, amp.imag, freq, width,
                                            phase, t_start, t_end, int(integrated))
        if rc:
            raise RuntimeError("orc_add_gaussian_source failed: %d" % rc)
        self.n_src += 1

    def add_cw_source(self, comp, lo, hi, amp, freq, width, t_start, t_end, slowness=3.0, integrated=True):
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        amp = complex(amp)
        rc = self.L.orc_add_cw_source(self.h, comp, _dp(lo), _dp(hi), amp.real, amp.imag, freq, width,
                                      t_start, t_end, slowness, int(integrated))
        if rc:
            raise RuntimeError("orc_add_cw_source failed: %d" % rc)
        self.n_src += 1

    def add_monitors(self, xyz, comp=0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.n_mon = xyz.shape[0]
        self.L.orc_add_monitors(self.h, comp, self.n_mon, _dp(xyz))

    def run(self, n_steps, save_span=1):
        return self.L.orc_run(self.h, int(n_steps), int(save_span))

    def step(self):
        self.L.orc_step(self.h)

    def monitors(self):
        ns = self.L.orc_n_saves(self.h)
        out = np.zeros((ns, self.n_mon, 2))
        if ns and self.n_mon:
            self.L.orc_read_monitors(self.h, _dp(out))
        return out

    def field(self, kind, comp, iset=0):
        p = self.L.orc_field(self.h, self.KIND[kind], comp, iset)
        return np.ctypeslib.as_array(p, shape=self.shape)

    def pol(self, isus, comp, iset=0):
        p = self.L.orc_pol(self.h, isus, comp, iset)
        return np.ctypeslib.as_array(p, shape=self.shape)

    def pml_sig(self, d):
        return np.ctypeslib.as_array(self.L.orc_pml_sig(self.h, d), shape=(2 * self.n[d] + 2,)).copy()

    def last_source_time(self):
        return self.L.orc_src_last_time(self.h)

    def dipole(self, isrc, t):
        out = np.zeros(2)
        self.L.orc_src_dipole(self.h, isrc, float(t), _dp(out))
        return complex(out[0], out[1])

    def src_info(self, isrc):
        lo = (C.c_int * 3)()
        hi = (C.c_int * 3)()
        self.L.orc_src_range(self.h, isrc, lo, hi)
        w = []
        for d in range(3):
            cnt = max(hi[d] - lo[d] + 1, 0)
            w.append(np.ctypeslib.as_array(self.L.orc_src_weights(self.h, isrc, d), shape=(max(cnt, 1),))[:cnt].copy())
        amp = np.zeros(2)
        self.L.orc_src_amp(self.h, isrc, _dp(amp))
        return list(lo), list(hi), w, complex(amp[0], amp[1])
