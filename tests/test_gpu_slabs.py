"""z-slabs inside the library (sj_connect_local / sj_connect_peers / sj_run_group): the boundary planes travel inside
the step kernels and the slabs order themselves on the device.  A stack of slabs must reproduce the single-slab run
bit for bit -- fields and monitors.  The one-device cases run on any GPU box (the slabs share the device and one
stream); the two-device cases need 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from sim_juncs_b200 import Sim
from sim_juncs_b200.materials import materials_from_regions
from sim_juncs_b200.parallel import SlabGroup

pytestmark = pytest.mark.gpu

N, A = (40, 36, 60), 6.0


def _slab(kz, device, prec="f64", nsets=2):
    shape = (N[2] + 1, N[1] + 1, N[0] + 1)
    k, j, i = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    masks = [((k > 31).astype(np.uint8) | ((i < 17) & (k > 20) & (k <= 31)).astype(np.uint8) << 1) for _ in range(3)]
    regs = (1.0, [2.25, 1.0], [[(1.1, 0.05, 1.3, 0)], [(1e-10, 0.04, 2.0e19, 1), (0.9, 0.1, 0.7, 0)]])
    g = Sim(N, A, pml=1.0, n_sets=nsets, device=device, precision=prec, kz=kz)
    g.set_materials(materials_from_regions(*regs), masks)
    g.add_gaussian_source(0, [0, 0, 2.0], [N[0] / A, N[1] / A, 2.0], 1.0, 0.4, 1.5, 0.3, 0.0, 18.0, True)
    # monitors on both sides of the cuts, one exactly between two planes that belong to different slabs
    g.add_monitors([[2.0, 1.6, 2.3], [1.0, 2.0, 3.0], [3.1, 2.2, 30.4 / A], [3.1, 2.2, 5.05], [2.2, 3.3, 6.7]], 0)
    return g


def _check(group, ref, steps):
    assert np.abs(ref.field(0, 0)).max() > 1e-3
    for c in range(6):
        for q in range(2):
            assert np.array_equal(group.field(c, q), ref.field(c, q)), (c, q)
    assert np.array_equal(group.monitors(), ref.monitors())


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("world", [2, 3])
def test_slabs_on_one_device_bitwise(prec, world):
    steps = 160
    ref = _slab(None, 0, prec)
    ref.run(steps, 4)
    group = SlabGroup(N[2] + 1, [0] * world, lambda kz, dev: _slab(kz, dev, prec))
    group.run(steps, 4)
    _check(group, ref, steps)
    # the cut planes carry signal (the pulse crossed them)
    cut = group.kz[0][1]
    assert np.abs(ref.field(0, 0)[cut - 1:cut + 1]).max() > 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slabs_on_two_devices_bitwise():
    steps = 160
    ref = _slab(None, 0)
    ref.run(steps, 4)
    group = SlabGroup(N[2] + 1, [0, 1], lambda kz, dev: _slab(kz, dev))
    group.run(steps, 4)
    _check(group, ref, steps)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, steps, out):
    import torch.distributed as dist
    from sim_juncs_b200.parallel import connect_slabs, slab_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kz = slab_range(N[2] + 1, rank, world)
    g = _slab(kz, rank)
    connect_slabs(g, rank, world)
    g.run(steps, 4)
    out[rank] = (kz, [g.field(c, q) for q in range(2) for c in range(6)], g.monitors())
    dist.barrier()
    g.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slabs_one_process_per_gpu_bitwise():
    """the bench / torchrun arrangement: CUDA-IPC handles over torch.distributed, then plain Sim.run on every rank"""
    import torch.multiprocessing as mp
    steps, world = 160, 2
    ref = _slab(None, 0)
    ref.run(steps, 4)
    ref_fields = [ref.field(c, q) for q in range(2) for c in range(6)]
    ref_mon = ref.monitors()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), steps, out), nprocs=world, join=True)
        res = dict(out)
    assert np.array_equal(sum(res[r][2] for r in range(world)), ref_mon)
    for r in range(world):
        (k0, k1), fields, _ = res[r]
        for a, b in zip(fields, ref_fields):
            assert np.array_equal(a, b[k0:k1])


def test_bound_geom_gpus_on_one_device(scene_json, tmp_path):
    """BoundGeom(gpus=[0, 0, 0]): the host-level entry to the slab stack (balanced cuts, run, save) against the plain
    single-GPU BoundGeom -- same series, same field_samples.h5 datasets"""
    from helpers import early_pulse, settings_from_doc
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.scene import Scene
    from sim_juncs_b200 import hdf5
    path = scene_json("Au_graphene_box")
    st = settings_from_doc(path)
    st.grid_num = 57
    st.resolution = 57 / 18.0
    st.post_source_t = 2.0
    st.save_span = 5
    one = BoundGeom(st, early_pulse(Scene.load(path), 0.2), n_sets=2, device=0)
    many = BoundGeom(st, early_pulse(Scene.load(path), 0.2), n_sets=2, gpus=[0, 0, 0])
    assert len(many.parts) == 3 and many.kz[0][0] == 0 and many.kz[-1][1] == st.grid_cells() + 1
    for bg in (one, many):
        bg.run()
    assert one.n_t_pts == many.n_t_pts and one.n_t_pts > 100
    a, b = np.array(one.field_times), np.array(many.field_times)
    assert np.abs(a).max() > 1e-6 and np.array_equal(a, b)
    f1 = hdf5.File(one.save_field_times(str(tmp_path / "one")))
    f2 = hdf5.File(many.save_field_times(str(tmp_path / "many")))
    cn = sorted(k for k in f1.keys() if k.startswith("cluster_"))[0]
    pn = sorted(k for k in f1[cn].keys() if k.startswith("point_"))[0]
    t1 = f1[cn][pn]["time"].read()
    t2 = f2[cn][pn]["time"].read()
    assert np.array_equal(t1["Re"], t2["Re"]) and np.array_equal(t1["Im"], t2["Im"])
    from helpers import rel_l2
    q1, q2 = f1[cn][pn]["frequency"].read(), f2[cn][pn]["frequency"].read()
    assert rel_l2(q2["Re"] + 1j * q2["Im"], q1["Re"] + 1j * q1["Im"]) < 1e-12      # device DFT vs host DFT of the same series
