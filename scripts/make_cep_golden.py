"""tests/golden/cep_signal.npz: what the reference's own `signal` class (/root/reference/scripts/phases.py) extracts from
a set of series -- run in the build container, where the reference is mounted (h5py and matplotlib are not installed
here and are stubbed: `signal` uses neither)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ("h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference/scripts")
import phases  # noqa: E402

phases.verbose = 0


def series_set():
    out = []
    rng = np.random.default_rng(1234)
    for k in range(12):                     # Gaussian pulses: carrier, width, arrival time, CEP, linear chirp
        n = int(rng.integers(300, 1600))
        dt = 0.05 + 0.1 * rng.random()
        t = np.arange(n) * dt
        t0 = t[-1] * (0.3 + 0.3 * rng.random())
        w = (2.0 + 4.0 * rng.random())
        f = (0.12 + 0.2 * rng.random()) / (4 * dt) * 0.4
        phi = rng.uniform(-np.pi, np.pi)
        chirp = rng.uniform(-0.002, 0.002)
        v = np.exp(-(t - t0) ** 2 / (2 * w * w)) * np.cos(2 * np.pi * f * (t - t0) + chirp * (t - t0) ** 2 + phi)
        out.append((t, v))
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_graphene_long.npz"))      # reference-driver series (Re part)
    tb = g["time_bounds"]
    ser = g["time"]
    t = np.arange(ser.shape[0]) * tb[2]
    for j in (0, 3, 6, 9):
        out.append((t, np.ascontiguousarray(ser[:, j].real)))
    return out


def main():
    keys = ["t0_guess", "f0", "f0_ind", "t0_corr", "phi_corr", "low_stren"]
    rows, ts, vs = [], [], []
    for t, v in series_set():
        n = len(v)
        guess = phases.signal._guess_t0_ind(types.SimpleNamespace(v_pts=v, v_abs=np.abs(v)))
        s = phases.signal(t, v, lowpass_inc=2.0)
        rows.append([guess, s.f0, s.f0_ind, s._t0_corr, s._phi_corr, s._low_stren])
        ts.append(t); vs.append(v)
    d = {"keys": np.array(keys), "rows": np.array(rows, dtype=float), "n": np.array([len(v) for v in vs])}
    for i, (t, v) in enumerate(zip(ts, vs)):
        d["t%d" % i] = t
        d["v%d" % i] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cep_signal.npz"), **d)
    print(np.array(rows))


if __name__ == "__main__":
    main()
