"""z-slab decomposition across GPUs: one process per GPU, one-plane halo exchange per half step.

SURVEY.md section 8(e): the path shards with a single nearest-neighbour exchange per half step --
after the H-pass the top owned Hx,Hy plane goes to rank+1 (it is that rank's lower halo); after the
E-pass the bottom owned Ex,Ey plane goes to rank-1 (its upper halo).  ADE, UPML auxiliaries, sources
and material lookups are cell-local.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) only
moves the planes; everything else is the single-GPU engine restricted to its slab.
"""
import torch
import torch.distributed as dist

HX, HY, EX, EY = 3, 4, 0, 1


def slab_range(n_planes, rank, world, weights=None):
    """Owned z-planes [k0, k1) of rank `rank`: contiguous, disjoint, covering [0, n_planes).
    `weights` (per plane cost, e.g. heavier PML planes) balances the cut; default is even."""
    if weights is None:
        per, rem = divmod(n_planes, world)
        k0 = rank * per + min(rank, rem)
        return k0, k0 + per + (1 if rank < rem else 0)
    tot = float(sum(weights))
    acc, cuts, target = 0.0, [0], 1
    for k, w in enumerate(weights):
        acc += w
        while target < world and acc >= tot * target / world:
            cuts.append(k + 1)
            target += 1
    while len(cuts) < world:
        cuts.append(n_planes)
    cuts.append(n_planes)
    for r in range(1, world + 1):          # every rank owns at least one plane
        cuts[r] = max(cuts[r], cuts[r - 1] + 1)
    cuts[world] = n_planes
    return cuts[rank], cuts[rank + 1]


class HaloExchanger:
    """Persistent P2P op lists for the two exchanges.  `plane(comp, set, k)` must return a contiguous
    tensor view of plane k (global index; k0-1 and k1 are the halo planes) of component `comp`."""

    def __init__(self, plane, kz, n_sets, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        k0, k1 = kz
        up, down = rank + 1, rank - 1
        self.ops_h, self.ops_e = [], []
        for q in range(n_sets):
            for c in (HX, HY):
                if up < world:
                    self.ops_h.append(dist.P2POp(dist.isend, plane(c, q, k1 - 1), up, group))
                if down >= 0:
                    self.ops_h.append(dist.P2POp(dist.irecv, plane(c, q, k0 - 1), down, group))
            for c in (EX, EY):
                if down >= 0:
                    self.ops_e.append(dist.P2POp(dist.isend, plane(c, q, k0), down, group))
                if up < world:
                    self.ops_e.append(dist.P2POp(dist.irecv, plane(c, q, k1), up, group))

    def _run(self, ops):
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def after_h(self):
        self._run(self.ops_h)

    def after_e(self):
        self._run(self.ops_e)

    def bytes_per_step(self):
        return sum(op.tensor.numel() * op.tensor.element_size() for op in self.ops_h + self.ops_e if op.op is dist.isend)


def cuda_plane_view(sim, device):
    """plane(comp, set, k) -> uint8 torch view of the engine's own device memory (no copies)."""
    cache = {}

    def plane(comp, q, k):
        key = (comp, q, k)
        if key not in cache:
            ptr, nbytes = sim.plane_ptr(comp, q, k)

            class _P:
                pass
            h = _P()
            h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            cache[key] = torch.as_tensor(h, device=device)
        return cache[key]
    return plane


class SlabRunner:
    """Drives one slab: sample / H-pass / exchange / E-pass / exchange / tick, all stream-ordered on
    torch's current CUDA stream so that NCCL and the kernels serialise without host syncs."""

    def __init__(self, sim, kz, n_sets, device, save_span=20):
        self.sim, self.kz, self.save_span = sim, kz, save_span
        self.stream = torch.cuda.current_stream(device).cuda_stream
        self.halo = HaloExchanger(cuda_plane_view(sim, device), kz, n_sets, dist.get_rank(), dist.get_world_size())
        self.i = 0

    def step(self):
        if self.i % self.save_span == 0:
            self.sim.sample(self.stream)
        self.sim.h_pass(self.kz[0], self.kz[1], self.stream)
        self.halo.after_h()
        self.sim.e_pass(self.kz[0], self.kz[1], self.stream)
        self.halo.after_e()
        self.sim.tick(self.stream)
        self.i += 1
