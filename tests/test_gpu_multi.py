"""N-GPU z-slabs over NCCL must reproduce the single-GPU run bit for bit (fields and monitors)."""
import os
import socket

import numpy as np
import pytest
import torch

from helpers import settings_from_doc

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _small_settings(scene_json):
    st = settings_from_doc(scene_json("Au_graphene_box"))
    st.grid_num = 73
    st.resolution = 73 / 18.0
    return st


def _worker(rank, world, port, scene_path, steps, overlap, out):
    import torch.distributed as dist
    from sim_juncs_b200.bound_geom import BoundGeom
    from sim_juncs_b200.parallel import SlabRunner, slab_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    class _S:
        pass
    import json
    st = _small_settings(lambda name: scene_path)
    n = st.grid_cells()
    kz = slab_range(n + 1, rank, world)
    from helpers import early_pulse
    from sim_juncs_b200.scene import Scene
    bg = BoundGeom(st, early_pulse(Scene.load(scene_path), 0.2), n_sets=2, kz=kz, device=rank)
    runner = SlabRunner(bg.sim, kz, 2, dev, save_span=5, overlap=overlap)
    for _ in range(steps):
        runner.step()
    torch.cuda.synchronize()
    bg.sim.sync()
    out[rank] = (kz, [bg.sim.field(c, q) for q in range(2) for c in range(6)], bg.sim.monitors())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", [False, True])
def test_two_gpu_slabs_bitwise(scene_json, overlap):
    import torch.multiprocessing as mp
    from sim_juncs_b200.bound_geom import BoundGeom
    from helpers import early_pulse
    from sim_juncs_b200.scene import Scene
    steps, world = 150, 2                          # the pulse (peak at step ~47) crosses the cut at plane 37
    path = scene_json("Au_graphene_box")
    st = _small_settings(scene_json)
    whole = BoundGeom(st, early_pulse(Scene.load(path), 0.2), n_sets=2, device=0)
    whole.sim.run(steps, 5)
    ref_fields = [whole.sim.field(c, q) for q in range(2) for c in range(6)]
    ref_mon = whole.sim.monitors()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), path, steps, overlap, out), nprocs=world, join=True)
        res = dict(out)
    mon = sum(res[r][2] for r in range(world))
    assert np.array_equal(mon, ref_mon)
    for r in range(world):
        (k0, k1), fields, _ = res[r]
        for a, b in zip(fields, ref_fields):
            assert np.array_equal(a, b[k0:k1])
    cut = res[0][0][1]
    assert np.abs(ref_fields[0][cut - 2:cut]).max() > 1e-6 and np.abs(ref_fields[0][cut:cut + 10]).max() > 1e-6
