"""Developer check: CUDA engine vs oracle on small cases (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import OracleSim
from sim_juncs_b200 import Sim
from sim_juncs_b200.materials import materials_from_regions

def rel(a, b):
    d = np.linalg.norm((a - b).ravel()); n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d

def case(n, a, pml, nsets, comp, lo, hi, mats, steps, prec, integrated=True, span=1):
    n = tuple(n)
    shape = (n[2] + 1, n[1] + 1, n[0] + 1)
    o = OracleSim(n, a, pml=pml, nsets=nsets)
    g = Sim(n, a, pml=pml, n_sets=nsets, precision=prec)
    if mats is not None:
        amb, reps, rpoles, masks = mats
        o.set_regions(amb, reps, rpoles, masks)
        g.set_materials(materials_from_regions(amb, reps, rpoles), masks)
    src = dict(freq=0.4, width=1.5, phase=0.3, t_start=1.0, t_end=1.0 + 12 * 1.5)
    o.add_gaussian_source(comp, lo, hi, 1.0, src["freq"], src["width"], src["phase"], src["t_start"], src["t_end"], integrated)
    g.add_gaussian_source(comp, lo, hi, 1.0, src["freq"], src["width"], src["phase"], src["t_start"], src["t_end"], integrated)
    L = [x / a for x in n]
    mon = [[L[0] / 2, L[1] / 2, L[2] / 2], [L[0] * 0.31, L[1] * 0.77, L[2] * 0.6], [L[0] / 3, L[1], L[2] / 3], [0.01, 0.02, 0.03]]
    o.add_monitors(mon, comp); g.add_monitors(mon, comp)
    t = time.time(); o.run(steps, span); to = time.time() - t
    t = time.time(); g.run(steps, span); tg = time.time() - t
    mo, mg = o.monitors(), g.monitors()
    worst = rel(mg[:, :, :nsets], mo[:, :, :nsets])
    # normalise every component by the largest component norm of its kind (symmetry-forbidden
    # components are pure round-off noise in both implementations)
    for kind, off in (("E", 0), ("H", 3)):
        for q in range(nsets):
            scale = max(np.linalg.norm(o.field(kind, c, q)) for c in range(3))
            for c in range(3):
                worst = max(worst, np.linalg.norm(g.field(off + c, q) - o.field(kind, c, q)) / scale)
    print("n=%s pml=%g sets=%d comp=%d prec=%s integ=%d mats=%s : worst rel-L2 %.3e  |mon| %.3e (oracle %.2fs gpu %.2fs)" % (
        n, pml, nsets, comp, prec, integrated, mats is not None, worst, np.abs(mo).max(), to, tg))
    return worst

def masks_for(n, a):
    shape = (n[2] + 1, n[1] + 1, n[0] + 1)
    k, j, i = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    ms = []
    for c in range(3):
        x = (i + 0.5 * (c == 0)) / a; y = (j + 0.5 * (c == 1)) / a; z = (k + 0.5 * (c == 2)) / a
        r0 = (z > 0.55 * n[2] / a)                       # substrate spanning the PML
        r1 = (x < 0.45 * n[0] / a) & (z > 0.3 * n[2] / a) & (z <= 0.62 * n[2] / a)   # lead, overlaps r0 a bit
        ms.append((r0.astype(np.uint8) | (r1.astype(np.uint8) << 1)))
    return ms

if __name__ == "__main__":
    ok = True
    for prec, tol in (("f64", 1e-11), ("f32", 2e-4)):
        w = case((20, 20, 20), 5.0, 1.0, 2, 1, [0, 0, 1], [4, 4, 1], None, 210, prec)
        ok &= w < tol
        w = case((20, 20, 20), 5.0, 1.0, 2, 1, [0, 0, 1], [4, 4, 1], None, 210, prec, integrated=False)
        ok &= w < tol
        n = (44, 36, 40); a = 8.0
        lor = (1.1, 0.05, 1.3, 0); dru = (1e-10, 0.04, 2.0e19, 1)
        mats = (1.0, [2.25, 1.0], [[lor], [dru, (0.9, 0.2, 0.7, 0)]], masks_for(n, a))
        w = case(n, a, 1.0, 2, 0, [0, 0, 1.0], [n[0] / a, n[1] / a, 1.0], mats, 300, prec, span=3)
        ok &= w < tol
        w = case(n, a, 0.0, 1, 2, [1.0, 1.0, 1.0], [2.0, 3.0, 2.2], mats, 120, prec)
        ok &= w < tol
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
