"""BoundGeom: host-side mirror of the reference's `bound_geom` (src/disp.hpp:142-200) on top of the
C ABI.  Same construction order, same unit conversions, same run-loop cadence; the meep objects
(structure, fields) are replaced by one sj_sim handle."""
import math
import os
import time

import numpy as np

from .engine import Sim
from .scene import COMPONENTS, LIGHT_SPEED, THICK_SCALE, Scene


class _SlabStack:
    """The `sim` of a BoundGeom that is spread over several GPUs: the z-slabs of one simulation, stepped together
    (sj_run_group); monitors are evaluated by the slab that owns them."""

    def __init__(self, sims):
        self.sims = sims
        self.dt = sims[0].dt
        self.n_sets = sims[0].n_sets

    def run(self, n_steps, save_span=1, sync=True):
        Sim.run_group(self.sims, n_steps, save_span, sync)

    def sync(self):
        for s in self.sims:
            s.sync()

    def monitors(self):
        return sum(s.monitors() for s in self.sims)

    def field(self, comp, iset=0):
        return np.concatenate([s.field(comp, iset) for s in self.sims], axis=0)

    def last_source_time(self):
        return self.sims[0].last_source_time()

    def launches(self):
        return sum(s.launches() for s in self.sims)


class BoundGeom:
    def __new__(cls, settings, scene=None, *args, gpus=None, **kw):
        if gpus is not None and gpus != 1 and gpus != [0] and cls is BoundGeom:
            return object.__new__(SlabBoundGeom)
        return object.__new__(cls)

    def __init__(self, settings, scene=None, precision="f64", n_sets=2, integrated=True, device=-1, kz=None,
                 verbose=False, phases=None, gpus=None):
        """settings: ParseSettings (after correct_defaults); scene: Scene, a path to a .geom file or to a
        scene JSON, or None to read settings.geom_fname as the reference does (disp.cpp:556).
        n_sets = 2 reproduces meep's complex fields (the reference never calls use_real_fields).
        phases = [phi_0, phi_1, ...] runs a CEP sweep as one batch instead (BASELINE config 5): one real
        field set per extra source phase, all sharing the rasterized materials; field_times is then
        indexed [phase][monitor]."""
        self.phases = None if phases is None else [float(x) for x in phases]
        if self.phases is not None:
            n_sets = len(self.phases)
        if scene is None:
            scene = settings.geom_fname
        if isinstance(scene, str):
            scene = Scene.load(scene) if scene.endswith(".json") else Scene.from_geom(scene, settings)
        if scene.ercode != 0:
            raise RuntimeError("Scene parsing failed (parse_ercode %d)" % scene.ercode)   # disp.cpp:562-565
        self.problem = scene
        self.settings = settings
        self.um_scale = settings.um_scale
        self.pml_thickness = settings.pml_thickness
        self.len = settings.len
        self.save_span = settings.save_span
        self.dump_raw = bool(settings.dump_raw)
        self.post_source_t = settings.post_source_t
        self.n_sets = n_sets
        if settings.n_dims != 3:
            raise NotImplementedError("only dimensions = 3 is supported (all shipped junction confs)")

        # ---- structure_from_settings (disp.cpp:482-550) ----
        n = settings.grid_cells()
        a = settings.resolution
        self.sim = Sim((n, n, n), a, pml=settings.pml_thickness, precision=precision, n_sets=n_sets, device=device,
                       kz=kz)
        nodes = scene.node_array()
        regions = []
        self.thicknesses = []
        for reg in scene.regions:
            thick = 1.0
            if reg.make_2d:
                thick = THICK_SCALE / settings.resolution
                M = nodes[reg.root].M          # rescale(): trans_mat = trans_mat * diag(1, 1, thick)
                for i in range(3):
                    M[3 * i + 0] = M[3 * i + 0] * 1.0
                    M[3 * i + 1] = M[3 * i + 1] * 1.0
                    M[3 * i + 2] = M[3 * i + 2] * thick
            self.thicknesses.append(thick)
            eps = reg.eps if reg.eps is not None else settings.ambient_eps      # add_region, disp.cpp:249-262
            poles = [(w0 / settings.um_scale, g / settings.um_scale, sg / thick, not use_denom)
                     for (w0, g, sg, use_denom) in reg.poles_raw]               # disp.cpp:539-541
            regions.append((reg.root, eps, poles))
        t0 = time.time()
        self.sim.rasterize(settings.ambient_eps, nodes, regions, settings.smooth_n, settings.smooth_rad)
        self.t_raster = time.time() - t0

        # ---- sources and monitors (disp.cpp:584-639) ----
        self.sources = []
        self.ttot = 0.0
        c_by_a = LIGHT_SPEED * settings.um_scale
        for info, (p1, p2) in zip(scene.sources, scene.source_boxes):
            lo = [min(p1[d], p2[d]) for d in range(3)]
            hi = [max(p1[d], p2[d]) for d in range(3)]
            frequency = 1 / (info.wavelen * settings.um_scale)
            width = info.width * c_by_a
            start_time = info.start_time * c_by_a
            end_time = info.end_time * c_by_a
            if info.component > 2:
                raise NotImplementedError("magnetic-current sources (%s) are not implemented" % COMPONENTS[info.component])
            if info.type == "gaussian":
                if verbose:
                    print("Adding Gaussian envelope: f=%f, w=%f, t_0=%f, t_f=%f (meep units)" % (frequency, width, start_time, end_time))
                self.sim.add_gaussian_source(info.component, lo, hi, info.amplitude, frequency, width, info.phase,
                                             start_time, end_time, integrated, set_phase=self.phases)
            else:
                # meep::continuous_src_time(frequency, width, start_time, end_time), disp.cpp:615-619 (slowness stays at
                # meep's default 3; the scene's `slowness` argument arrives here as the width, disp.cpp:343-347)
                if verbose:
                    print("Adding continuous wave: f=%f, w=%f, t_0=%f, t_f=%f (meep units)" % (frequency, width, start_time, end_time))
                if not math.isfinite(end_time) or end_time > 1e300:
                    raise ValueError("CW_source without end_time never ends: the reference's run() would not terminate either")
                self.sim.add_cw_source(info.component, lo, hi, info.amplitude, frequency, width, start_time, end_time,
                                       3.0, integrated, set_phase=self.phases)
            self.sources.append(info)
            self.ttot = self.sim.last_source_time() + self.post_source_t * LIGHT_SPEED * settings.um_scale
        self.monitor_locs = [tuple(p) for p in scene.monitor_locs]
        self.monitor_clusters = list(scene.monitor_clusters)
        if self.monitor_locs:
            self.sim.add_monitors(np.array(self.monitor_locs), comp=0)      # always Ex (disp.cpp:724)
        self.field_times = []
        self.n_t_pts = 0

    # ---- getters (disp.hpp:146-149) ----
    def get_monitor_locs(self):
        return self.monitor_locs

    def get_field_times(self):
        return self.field_times

    def get_sources(self):
        return self.sources

    def get_n_monitor_clusters(self):
        return len(self.monitor_clusters)

    def fs_to_meep_time(self, t):
        return t * LIGHT_SPEED * self.um_scale

    def meep_time_to_fs(self, t):
        return t / (LIGHT_SPEED * self.um_scale)

    # ---- run (disp.cpp:690-749) ----
    def run(self, fname_prefix=None, verbose=False):
        if self.save_span == 0:
            self.save_span = 1
        dt = self.sim.dt
        self.n_t_pts = int((self.ttot + dt / 2) / dt)
        t0 = time.time()
        if fname_prefix is not None:                 # fields.output_hdf5(meep::Dielectric, ...), disp.cpp:696
            from .output import write_eps_h5
            write_eps_h5(self, fname_prefix)
        try:
            if self.dump_raw and fname_prefix is not None:
                # disp.cpp:719-741 with dump_raw: at every save point the whole Ex field goes to ex-<time>.h5 first
                from .output import write_ex_h5
                done = 0
                while done < self.n_t_pts:
                    write_ex_h5(self, fname_prefix, done * dt)
                    chunk = min(self.save_span, self.n_t_pts - done)
                    self.sim.run(chunk, self.save_span, sync=False)
                    done += chunk
            else:
                self.sim.run(self.n_t_pts, self.save_span, sync=False)
            self.sim.sync()
        except Exception as err:                     # divergence is reported, the save still happens
            print("error on step %d: %s" % (self.n_t_pts, err))
        self.t_run = time.time() - t0
        series = self.sim.monitors()                 # [n_saves][n_mon][n_sets]
        if self.phases is not None:
            self.field_times = [[series[:, j, q].astype(np.complex128) for j in range(series.shape[1])]
                                for q in range(series.shape[2])]
            return
        if self.n_sets >= 2:
            cplx = series[:, :, 0] + 1j * series[:, :, 1]
        else:
            cplx = series[:, :, 0].astype(np.complex128)
        self.field_times = [cplx[:, j].copy() for j in range(cplx.shape[1])]
        if verbose:
            print("Simulations completed")

    def time_bounds(self):
        """/info/time_bounds (disp.cpp:817-825)."""
        ttot_fs = self.meep_time_to_fs(self.ttot)
        return [0.0, ttot_fs, ttot_fs * self.save_span / self.n_t_pts if self.n_t_pts else 0.0]

    def save_field_times(self, fname_prefix):
        """disp.cpp:758-923: <prefix>/field_samples.h5 through the package's own HDF5 encoder (no libhdf5 here),
        plus <prefix>/field_samples.npz with the HDF5 paths as keys.  Returns the .h5 path."""
        from .output import save_field_samples
        return save_field_samples(self, fname_prefix)


class SlabBoundGeom(BoundGeom):
    """BoundGeom(..., gpus=N) or gpus=[device, ...]: the same simulation cut into z-slabs of equal bytes per step, one per
    GPU, inside this process (the reference's counterpart is meep under MPI, src/main.cpp:20).  Results equal the single-
    GPU run bit for bit; run() / save_field_times() / the getters are BoundGeom's."""

    def __init__(self, settings, scene=None, precision="f64", n_sets=2, integrated=True, device=-1, kz=None,
                 verbose=False, phases=None, gpus=None):
        from .parallel import slab_range
        devices = list(range(gpus)) if isinstance(gpus, int) else list(gpus)
        world = len(devices)
        if scene is None:
            scene = settings.geom_fname
        if isinstance(scene, str):
            scene = Scene.load(scene) if scene.endswith(".json") else Scene.from_geom(scene, settings)
        n_planes = settings.grid_cells() + 1

        def build(cuts):
            return [BoundGeom(settings, scene, precision=precision, n_sets=n_sets, integrated=integrated, device=devices[r],
                              kz=cuts[r], verbose=verbose and r == 0, phases=phases) for r in range(world)]
        parts = build([slab_range(n_planes, r, world) for r in range(world)])
        weights = [float(x) for p in parts for x in p.sim.plane_costs()]
        cuts = [slab_range(n_planes, r, world, weights) for r in range(world)]
        for p in parts:
            p.sim.close()
        self.parts = build(cuts)
        self.kz = cuts
        for lo, up in zip(self.parts[:-1], self.parts[1:]):
            Sim.connect_local(lo.sim, up.sim)
        first = self.parts[0]
        for k, v in first.__dict__.items():           # scene, settings, sources, monitors, ttot ... are the same in every part
            if k != "sim":
                setattr(self, k, v)
        self.sim = _SlabStack([p.sim for p in self.parts])
        self.t_raster = max(p.t_raster for p in self.parts)
