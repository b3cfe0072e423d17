"""oracle/cep_oracle.py (NumPy restatement of the reference's phases.signal front half) against what the reference class
itself extracted (tests/golden/cep_signal.npz, written by scripts/make_cep_golden.py from /root/reference/scripts/phases.py)."""
import os

import numpy as np

from oracle import cep_oracle as co


def _cases(golden):
    g = np.load(os.path.join(golden, "cep_signal.npz"))
    keys = [str(k) for k in g["keys"]]
    for i, row in enumerate(g["rows"]):
        yield g["t%d" % i], g["v%d" % i], dict(zip(keys, row))


def test_restatement_equals_reference_class(golden):
    n = 0
    for t, v, ref in _cases(golden):
        assert ref["low_stren"] == 0.0          # none of the golden series takes the low-pass branch
        got = co.signal_params(t, v)
        assert got["t0_ind"] == int(ref["t0_guess"])
        assert got["f0_ind"] == int(ref["f0_ind"])
        assert abs(got["f0"] - ref["f0"]) <= 1e-12 * abs(ref["f0"])
        assert abs(got["t0_corr"] - ref["t0_corr"]) <= 1e-9 * max(1.0, abs(ref["t0_corr"]))
        assert abs(got["phi_corr"] - ref["phi_corr"]) <= 1e-9
        n += 1
    assert n >= 16


def test_fix_angle_range():
    for th in np.linspace(-20, 20, 401):
        a = co.fix_angle(th)
        assert -np.pi < a <= np.pi + 1e-15
        assert abs(np.exp(1j * a) - np.exp(1j * th)) < 1e-12
