"""The TMA-staged column kernels (csrc/sj_tma.cuh) against the register kernels (csrc/sj_kernels.cuh): same arithmetic,
different data movement -- the fields must agree bit for bit, whatever the tile shapes and material classes.  Both
paths are then pinned to the oracle by test_gpu_parity.py (which runs the default path = TMA)."""
import os

import numpy as np
import pytest

from sim_juncs_b200 import Sim
from sim_juncs_b200.materials import materials_from_regions

pytestmark = pytest.mark.gpu


def _sim(mode, prec, n, nsets, pml, src_z):
    old = os.environ.get("SJ_TMA")
    os.environ["SJ_TMA"] = str(mode)         # read by sj_create
    try:
        a = 6.0
        shape = (n[2] + 1, n[1] + 1, n[0] + 1)
        k, j, i = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        # region 0: one Lorentz pole above plane 22; region 1: Drude + Lorentz block that also crosses the x-low PML
        masks = [((k > 22).astype(np.uint8) | ((i < 17) & (k > 14) & (k <= 22)).astype(np.uint8) << 1) for _ in range(3)]
        regs = (1.0, [2.25, 1.0], [[(1.1, 0.05, 1.3, 0)], [(1e-10, 0.04, 2.0e19, 1), (0.9, 0.1, 0.7, 0)]])
        g = Sim(n, a, pml=pml, n_sets=nsets, device=0, precision=prec)
        g.set_materials(materials_from_regions(*regs), masks)
        g.add_gaussian_source(0, [0, 0, src_z], [n[0] / a, n[1] / a, src_z], 1.0, 0.4, 1.5, 0.3, 0.0, 18.0, True)
        g.add_monitors([[2.0, 1.6, 2.3], [1.0, 2.0, 3.0]], 0)
        return g
    finally:
        if old is None:
            os.environ.pop("SJ_TMA", None)
        else:
            os.environ["SJ_TMA"] = old


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("src_z", [1.5, 0.5])        # source plane in the interior / inside the z-low PML
def test_tma_equals_register_kernels_bitwise(prec, src_z):
    n = (40, 36, 44)
    ref = _sim(0, prec, n, 2, 1.0, src_z)
    ref.run(120, 4)
    assert np.abs(ref.field(0, 0)).max() > 1e-3
    for mode in (1, 2, 3):                            # H-pass only, E-pass only, both through TMA
        g = _sim(mode, prec, n, 2, 1.0, src_z)
        g.run(120, 4)
        for c in range(6):
            for q in range(2):
                assert np.array_equal(ref.field(c, q), g.field(c, q)), (mode, c, q)
        assert np.array_equal(ref.monitors(), g.monitors())
        g.close()
    ref.close()


def test_tma_no_pml_and_odd_sizes():
    """no absorber (interior tiles only, metallic wall handled by the interior box) and sizes that leave ragged tiles"""
    n = (37, 29, 31)
    ref = _sim(0, "f64", n, 1, 0.0, 1.5)
    g = _sim(3, "f64", n, 1, 0.0, 1.5)
    for s in (ref, g):
        s.run(90, 3)
    for c in range(6):
        assert np.array_equal(ref.field(c, 0), g.field(c, 0)), c


def _sim_env(env, prec, n):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})      # read when the work lists are built (sj_create / sj_set_materials)
    try:
        return _sim(3, prec, n, 2, 1.0, 1.5)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_fused_wavefront_and_tall_tiles_bitwise(prec):
    """One launch per step (H-pass and E-pass items in one queue, z wavefront with per-chunk dependencies) against two
    launches per step, tall tiles (two rows per thread) against plain ones, several chunk lengths: identical fields."""
    n = (52, 47, 44)
    ref = _sim_env({"SJ_TMA_WAVE": 0, "SJ_TMA_SUB": 1}, prec, n)
    ref.run(90, 4)
    l_ref = ref.launches()
    assert np.abs(ref.field(0, 0)).max() > 1e-3
    for env in ({"SJ_TMA_WAVE": 0, "SJ_TMA_SUB": 2}, {"SJ_TMA_WAVE": 2, "SJ_TMA_LEAD": 1}, {"SJ_TMA_WAVE": 3, "SJ_TMA_LEAD": 2},
                {"SJ_TMA_WAVE": 8}, {"SJ_TMA_WAVE": 64}, {}):
        g = _sim_env(env, prec, n)
        g.run(90, 4)
        for c in range(6):
            for q in range(2):
                assert np.array_equal(ref.field(c, q), g.field(c, q)), (env, c, q)
        assert np.array_equal(ref.monitors(), g.monitors())
        if env.get("SJ_TMA_WAVE", 8) != 0:
            assert g.launches() < l_ref, "the fused step is one launch"      # (plus the monitor samples)
        g.close()
    ref.close()
