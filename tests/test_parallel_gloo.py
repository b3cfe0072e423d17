"""Host-side slab logic under a real 2-rank process group (gloo, CPU): partition properties and the
halo exchange schedule (who sends which plane to whom, after which half-pass)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sim_juncs_b200.parallel import EX, EY, HX, HY, HaloExchanger, slab_range


@pytest.mark.parametrize("n,world", [(182, 1), (182, 2), (182, 8), (1010, 8), (9, 8), (543, 4)])
def test_slab_partition_covers_exactly(n, world):
    cuts = [slab_range(n, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    for a, b in zip(cuts, cuts[1:]):
        assert a[1] == b[0]
    sizes = [k1 - k0 for k0, k1 in cuts]
    assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


def test_weighted_partition_balances_cost():
    n, world = 218, 4
    w = np.ones(n)
    w[:13] = w[-13:] = 1.7            # PML planes cost more bytes per cell
    cuts = [slab_range(n, r, world, list(w)) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    cost = [w[k0:k1].sum() for k0, k1 in cuts]
    assert max(cost) / min(cost) < 1.1
    assert cuts[0][1] - cuts[0][0] < cuts[1][1] - cuts[1][0]      # edge slabs are thinner


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_planes, n_sets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k0, k1 = slab_range(n_planes, rank, world)
    # local storage: planes k0-1 .. k1 for 6 components; value encodes (comp, set, global plane, owner)
    store = {}

    def plane(c, q, k):
        if (c, q, k) not in store:
            store[(c, q, k)] = torch.full((12,), -1.0)
        return store[(c, q, k)]
    for c in range(6):
        for q in range(n_sets):
            for k in range(k0, k1):
                plane(c, q, k).fill_(1000 * c + 100 * q + k + 0.5 * rank)
    ex = HaloExchanger(plane, (k0, k1), n_sets, rank, world)
    ex.after_h()
    ok = True
    for q in range(n_sets):
        for c in (HX, HY):
            if rank > 0:     # lower halo came from rank-1's top owned plane
                ok &= bool((plane(c, q, k0 - 1) == 1000 * c + 100 * q + (k0 - 1) + 0.5 * (rank - 1)).all())
        for c in (EX, EY):   # E halos must still be untouched
            if rank < world - 1:
                ok &= bool((plane(c, q, k1) == -1.0).all())
    ex.after_e()
    for q in range(n_sets):
        for c in (EX, EY):
            if rank < world - 1:   # upper halo came from rank+1's bottom owned plane
                ok &= bool((plane(c, q, k1) == 1000 * c + 100 * q + k1 + 0.5 * (rank + 1)).all())
    n_send = (2 * n_sets if rank < world - 1 else 0) + (2 * n_sets if rank > 0 else 0)
    ok &= ex.bytes_per_step() == n_send * 12 * 4
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), 19, 2, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_halo_exchange_world3_gloo():
    world = 3
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), 10, 1, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True, 2: True}


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("overlap", [False, True])
def test_step_schedule_orders_boundary_planes_before_their_exchange(world, overlap):
    from sim_juncs_b200.parallel import step_schedule
    n = 37
    for rank in range(world):
        kz = slab_range(n, rank, world)
        acts = step_schedule(kz, rank, world, overlap)
        for which in "he":
            spans = [(a[1], a[2]) for a in acts if a[0] == which]
            covered = sorted(k for a, b in spans for k in range(a, b))
            assert covered == list(range(kz[0], kz[1]))                      # every owned plane exactly once
            post, wait = acts.index(("post_" + which,)), acts.index(("wait_" + which,))
            assert post < wait
            # the plane the neighbour receives is finished before the exchange is posted
            need = kz[1] - 1 if which == "h" else kz[0]
            has_peer = (rank + 1 < world) if which == "h" else (rank > 0)
            if has_peer:
                done = [k for a in acts[:post] if a[0] == which for k in range(a[1], a[2])]
                assert need in done
                if overlap and kz[1] - kz[0] >= 2:
                    assert len(done) == 1                                    # ... and only that plane
        # the E-pass starts after the H halo has arrived; the step ends with the E halo in place
        first_e = min(i for i, a in enumerate(acts) if a[0] == "e")
        assert acts.index(("wait_h",)) < first_e and acts[-1] == ("wait_e",)
