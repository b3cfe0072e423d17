"""Small cases for compute-sanitizer (scripts/sanitize.sh): the reference's 20^3 unit-test grid (C1) and a dispersive + PML
box with two slabs on one device (boundary-plane exchange included), a few steps each."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sim_juncs_b200 import Sim  # noqa: E402
from sim_juncs_b200.materials import materials_from_regions  # noqa: E402
from sim_juncs_b200.parallel import SlabGroup  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
# C1: 20^3, a = 5, pml 1.0 (5 cells), vacuum, Ey plane source, 2 monitors
g = Sim((20, 20, 20), 5.0, pml=1.0, n_sets=2, device=0)
g.add_gaussian_source(1, [0, 0, 1.0], [4.0, 4.0, 1.0], 1.0, 0.3759, 1.199, 0.0, 0.0, 20.0, True)
g.add_monitors([[2.0, 2.0, 2.0], [1.0, 3.0, 2.5]], 0)
g.run(steps, 1)
print("C1 max|Ey|", np.abs(g.field(1, 0)).max())
g.close()


def slab(kz, dev):
    n, a = (24, 20, 28), 6.0
    shape = (n[2] + 1, n[1] + 1, n[0] + 1)
    k, j, i = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    masks = [((k > 16).astype(np.uint8) | ((i < 9) & (k > 10) & (k <= 16)).astype(np.uint8) << 1) for _ in range(3)]
    regs = (1.0, [2.25, 1.0], [[(1.1, 0.05, 1.3, 0)], [(1e-10, 0.04, 2.0e19, 1), (0.9, 0.1, 0.7, 0)]])
    s = Sim(n, a, pml=1.0, n_sets=2, device=dev, kz=kz)
    s.set_materials(materials_from_regions(*regs), masks)
    s.add_gaussian_source(0, [0, 0, 1.0], [n[0] / a, n[1] / a, 1.0], 1.0, 0.4, 1.5, 0.3, 0.0, 19.0, True)
    s.add_monitors([[2.0, 1.6, 2.3]], 0)
    return s


grp = SlabGroup(29, [0, 0], slab)
grp.run(steps, 2)
print("slabs max|Ex|", np.abs(grp.field(0, 0)).max(), "monitor", grp.monitors()[-1].ravel())
print(g.extract_cep(0.1)["status"] if False else "done")
