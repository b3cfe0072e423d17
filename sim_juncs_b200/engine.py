"""Object wrapper over the C ABI (one `Sim` = one sj_sim handle = one GPU, one z-slab)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SjCsgNode, SjGrid, SjMaterial, SjPole, SjRegion


class SjError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_pole(omega0, gamma, sigma, drude):
    p = SjPole()
    p.omega0, p.gamma, p.sigma, p.drude = float(omega0), float(gamma), float(sigma), int(bool(drude))
    return p


class Sim:
    def __init__(self, n, a, pml=0.0, courant=0.5, R=1e-15, precision="f64", n_sets=1, kz=None, device=-1):
        self.L = _lib.load()
        g = SjGrid()
        g.n[:] = [int(x) for x in n]
        g.a, g.courant, g.pml_thickness, g.pml_R = float(a), float(courant), float(pml), float(R)
        g.precision = _lib.SJ_F32 if precision in ("f32", "fp32", 1) else _lib.SJ_F64
        g.n_sets = int(n_sets)
        g.kz0, g.kz1 = (0, 0) if kz is None else (int(kz[0]), int(kz[1]))
        g.device = int(device)
        self.n = tuple(int(x) for x in n)
        self.n_sets = int(n_sets)
        self.kz = (0, self.n[2] + 1) if kz is None else (int(kz[0]), int(kz[1]))
        self.h = C.c_void_p()
        rc = self.L.sj_create(C.byref(g), C.byref(self.h))
        if rc:
            msg = self.L.sj_last_error(self.h if self.h else None)
            if self.h:
                self.L.sj_destroy(self.h)
                self.h = None
            raise SjError("sj_create failed (%d): %s" % (rc, msg.decode() if msg else ""))
        self.n_mon = 0
        self.shape = (self.kz[1] - self.kz[0], self.n[1] + 1, self.n[0] + 1)

    def close(self):
        if getattr(self, "h", None):
            self.L.sj_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc:
            raise SjError("libsimjuncs_b200 error %d: %s" % (rc, self.L.sj_last_error(self.h).decode()))

    @property
    def dt(self):
        return self.L.sj_dt(self.h)

    # ---- materials ----
    def set_materials(self, mats, idx=None):
        """mats: list of (eps_inf, [(omega0,gamma,sigma,drude),...]); idx: 3 uint8 arrays over the GLOBAL grid."""
        arr = (SjMaterial * len(mats))()
        for m, (eps, poles) in enumerate(mats):
            arr[m].eps_inf = float(eps)
            arr[m].n_poles = len(poles)
            for q, p in enumerate(poles):
                arr[m].poles[q] = make_pole(*p)
        ptrs = [None, None, None]
        keep = []
        if idx is not None:
            gshape = (self.n[2] + 1, self.n[1] + 1, self.n[0] + 1)
            for c in range(3):
                a = np.ascontiguousarray(idx[c], dtype=np.uint8).reshape(gshape)
                keep.append(a)
                ptrs[c] = a.ctypes.data_as(C.POINTER(C.c_uint8))
        self._ck(self.L.sj_set_materials(self.h, len(mats), arr, ptrs[0], ptrs[1], ptrs[2]))

    def rasterize(self, ambient_eps, nodes, regions, smooth_n=0, smooth_rad=0.0):
        """nodes: ctypes array of SjCsgNode; regions: list of (root, eps, poles); smooth_n / smooth_rad: the reference's
        stochastic boundary smoothing (params.conf, disp.cpp:56-112)."""
        reg = (SjRegion * max(len(regions), 1))()
        for r, (root, eps, poles) in enumerate(regions):
            reg[r].root = int(root)
            reg[r].eps = float(eps)
            reg[r].n_poles = len(poles)
            for q, p in enumerate(poles):
                reg[r].poles[q] = make_pole(*p)
        self._ck(self.L.sj_rasterize_smooth(self.h, float(ambient_eps), len(nodes), nodes, len(regions), reg,
                                            int(smooth_n), float(smooth_rad)))

    def material_ids(self, comp):
        """index into material_table() at every owned Yee point of E component comp"""
        out = np.zeros(self.shape, dtype=np.uint8)
        self._ck(self.L.sj_get_material_ids(self.h, comp, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def region_masks(self, comp):
        out = np.zeros(self.shape, dtype=np.uint8)
        self._ck(self.L.sj_get_region_masks(self.h, comp, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def material_table(self):
        n = C.c_int32()
        arr = (SjMaterial * 256)()
        self._ck(self.L.sj_get_material_table(self.h, C.byref(n), arr, 256))
        return [(arr[i].eps_inf, [(arr[i].poles[q].omega0, arr[i].poles[q].gamma, arr[i].poles[q].sigma,
                                   arr[i].poles[q].drude) for q in range(arr[i].n_poles)]) for i in range(n.value)]

    # ---- sources / monitors ----
    def add_gaussian_source(self, comp, lo, hi, amp, freq, width, phase, t_start, t_end, integrated=True,
                            set_phase=None):
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        amp = complex(amp)
        sp = None
        if set_phase is not None:
            spa = np.ascontiguousarray(set_phase, dtype=np.float64)
            assert spa.size == self.n_sets
            sp = _dp(spa)
        self._ck(self.L.sj_add_gaussian_source(self.h, comp, _dp(lo), _dp(hi), amp.real, amp.imag, freq, width, phase,
                                               t_start, t_end, int(integrated), sp))

    def add_cw_source(self, comp, lo, hi, amp, freq, width, t_start, t_end, slowness=3.0, integrated=True, set_phase=None):
        """meep::continuous_src_time over a volume (reference CW_source, src/disp.cpp:615-619)."""
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        amp = complex(amp)
        sp = None
        if set_phase is not None:
            spa = np.ascontiguousarray(set_phase, dtype=np.float64)
            assert spa.size == self.n_sets
            sp = _dp(spa)
        self._ck(self.L.sj_add_cw_source(self.h, comp, _dp(lo), _dp(hi), amp.real, amp.imag, freq, width, t_start, t_end,
                                         float(slowness), int(integrated), sp))

    def add_custom_source(self, comp, lo, hi, amp, dipole, last_time, integrated=True, set_phase=None):
        """dipole: callable t -> complex, the waveform of any meep::src_time subclass (sj_add_custom_source)."""
        from ._lib import DIPOLE_FN
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        amp = complex(amp)

        def tramp(ctx, t, out):
            v = complex(dipole(t))
            out[0] = v.real
            out[1] = v.imag
        cb = DIPOLE_FN(tramp)
        self._callbacks = getattr(self, "_callbacks", []) + [cb]      # must outlive the simulation
        ph = None if set_phase is None else np.ascontiguousarray(set_phase, dtype=np.float64)
        self._ck(self.L.sj_add_custom_source(self.h, comp, _dp(lo), _dp(hi), amp.real, amp.imag, cb, None, float(last_time),
                                             int(integrated), None if ph is None else _dp(ph)))

    def sample_at(self, xyz, comp=0):
        """fields.get_field at arbitrary points, now: [n, n_sets]"""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((len(xyz), self.n_sets))
        self._ck(self.L.sj_sample_at(self.h, comp, len(xyz), _dp(xyz), _dp(out)))
        return out

    def last_source_time(self):
        return self.L.sj_last_source_time(self.h)

    def add_monitors(self, xyz, comp=0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.n_mon = xyz.shape[0]
        self._ck(self.L.sj_add_monitors(self.h, comp, self.n_mon, _dp(xyz)))

    # ---- stepping ----
    def run(self, n_steps, save_span=1, sync=True):
        self._ck(self.L.sj_run(self.h, int(n_steps), int(save_span)))
        if sync:
            self.sync()

    def sync(self):
        self._ck(self.L.sj_sync(self.h))

    def monitors(self):
        ns = self.L.sj_n_samples(self.h)
        out = np.zeros((ns, self.n_mon, self.n_sets))
        if ns and self.n_mon:
            self._ck(self.L.sj_read_monitors(self.h, _dp(out)))
        return out

    def spectra(self, set_re=0, set_im=None):
        """[n_monitors][T] complex: the reference's `frequency` transform of every monitor series, on the device."""
        t = C.c_int32()
        im = -1 if set_im is None else int(set_im)
        self._ck(self.L.sj_read_spectra(self.h, int(set_re), im, C.byref(t), None))
        out = np.zeros((self.n_mon, t.value, 2))
        if out.size:
            self._ck(self.L.sj_read_spectra(self.h, int(set_re), im, C.byref(t), _dp(out)))
        return out[:, :, 0] + 1j * out[:, :, 1]

    def extract_cep(self, dt_sample):
        """pulse parameters of every monitor series on the device (sj_extract_cep): dict of [n_sets, n_mon] arrays"""
        out = np.zeros((self.n_sets, self.n_mon, 10))
        self._ck(self.L.sj_extract_cep(self.h, float(dt_sample), _dp(out)))
        band = out[:, :, 8]
        return dict(f0=out[:, :, 0], f0_ind=out[:, :, 1].astype(int), t0_ind=out[:, :, 2].astype(int), t0_corr=out[:, :, 3],
                    phi_corr=out[:, :, 4], slope=out[:, :, 5], intercept=out[:, :, 6], rvalue=out[:, :, 7],
                    f_min=(band % 65536).astype(int), f_max=(band // 65536).astype(int), status=out[:, :, 9].astype(int))

    def field(self, comp, iset=0):
        out = np.zeros(self.shape)
        self._ck(self.L.sj_get_field(self.h, comp, iset, _dp(out)))
        return out

    def h_pass(self, k0, k1, stream=None):
        self._ck(self.L.sj_pass(self.h, 0, k0, k1, stream))

    def e_pass(self, k0, k1, stream=None):
        self._ck(self.L.sj_pass(self.h, 1, k0, k1, stream))

    def tick(self, stream=None):
        self._ck(self.L.sj_tick(self.h, stream))

    def sample(self, stream=None):
        self._ck(self.L.sj_sample(self.h, stream))

    def plane_ptr(self, comp, iset, k):
        p = C.c_void_p()
        nb = C.c_size_t()
        self._ck(self.L.sj_plane_ptr(self.h, comp, iset, k, C.byref(p), C.byref(nb)))
        return p.value, nb.value

    @staticmethod
    def halo_exchange(lower, upper, which):
        rc = lower.L.sj_halo_exchange(lower.h, upper.h, int(which))
        if rc:
            raise SjError("sj_halo_exchange failed %d: %s" % (rc, lower.L.sj_last_error(lower.h).decode()))

    # ---- z-slabs inside the library (include/sim_juncs_b200.h, "z-slabs inside the library") ----
    def export_peer(self):
        """256-byte CUDA-IPC description of this slab, to be handed to the neighbouring slabs' processes"""
        buf = C.create_string_buffer(256)
        self._ck(self.L.sj_export_peer(self.h, buf))
        return buf.raw

    def connect_peers(self, lower=None, upper=None):
        """lower / upper: export_peer() bytes of the slab below / above (other processes), None where there is none"""
        self._ck(self.L.sj_connect_peers(self.h, lower, upper))

    @staticmethod
    def connect_local(lower, upper):
        """two stacked slabs held by this process (any devices with peer access, or the same device)"""
        rc = lower.L.sj_connect_local(lower.h, upper.h)
        if rc:
            raise SjError("sj_connect_local failed %d: %s" % (rc, lower.L.sj_last_error(lower.h).decode()))

    @staticmethod
    def run_group(sims, n_steps, save_span=1, sync=True):
        """sj_run for all slabs of this process, bottom to top"""
        arr = (C.c_void_p * len(sims))(*[s.h for s in sims])
        rc = sims[0].L.sj_run_group(arr, len(sims), int(n_steps), int(save_span))
        if rc:
            raise SjError("sj_run_group failed %d: %s" % (rc, "; ".join(s.L.sj_last_error(s.h).decode() for s in sims)))
        if sync:
            for s in sims:
                s.sync()

    def run_timed(self, n_steps, save_span=1):
        ms = C.c_double()
        self._ck(self.L.sj_run_timed(self.h, int(n_steps), int(save_span), C.byref(ms)))
        return ms.value

    def profile_kernels(self, reps=5):
        out = (C.c_double * 4)()
        self._ck(self.L.sj_profile_kernels(self.h, int(reps), out))
        return dict(h_interior=out[0], e_interior=out[1], h_pml=out[2], e_pml=out[3])

    def counts(self):
        out = (C.c_double * 6)()
        self._ck(self.L.sj_get_counts(self.h, out))
        return dict(cells=out[0], interior_cells=out[1], pml_kernel_cells=out[2], pole_points=out[3],
                    pole_points_interior=out[4], pml_cells=out[5])

    def launches(self):
        n = C.c_int64()
        self.L.sj_get_stats(self.h, C.byref(n), None)
        return n.value

    def plane_costs(self):
        """algorithmic bytes per step of every owned plane (sj_plane_costs)"""
        out = np.zeros(self.kz[1] - self.kz[0])
        self._ck(self.L.sj_plane_costs(self.h, _dp(out)))
        return out

    def memory(self):
        """device bytes of this slab by kind (sj_memory)"""
        out = (C.c_double * 6)()
        self._ck(self.L.sj_memory(self.h, out))
        return dict(fields=out[0], polarisation=out[1], pml=out[2], materials=out[3], total=out[4], pol_planes=int(out[5]))

    def h2d_bytes(self):
        """bytes of source drive table uploaded to the device so far"""
        b = C.c_double()
        self.L.sj_get_stats(self.h, None, C.byref(b))
        return b.value

    def bytes_per_step(self):
        return self.L.sj_bytes_per_step(self.h)

    def steps_done(self):
        return self.L.sj_steps_done(self.h)
