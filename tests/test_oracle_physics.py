"""Physics pins of the FDTD oracle (the reference holds no golden field value).

Cavity resonances are used because the metallic outer wall makes them exact statements:
  * vacuum PEC cavity: discrete Yee dispersion  sin^2(w dt/2)/dt^2 = sum_i sin^2(k_i dx/2)/dx^2
  * cavity filled with a Lorentz / Drude medium: eps(w) w^2 = w_vac^2 with the reference's own
    permittivity model (scripts/utils.py:18-50: eps = eps_inf + sum sigma f0^2 / (f0^2 - f^2 - i f gamma))
  * PML: the pulse leaves the box (residual << peak); lossy medium: energy decays monotonically.
"""
import numpy as np
import pytest

from oracle.oracle import OracleSim


def _spectrum_peaks(sig, dt, fmin, fmax, pad=8):
    n = len(sig)
    w = np.hanning(n)
    spec = np.abs(np.fft.rfft(sig * w, n * pad))
    f = np.fft.rfftfreq(n * pad, dt)
    sel = (f > fmin) & (f < fmax)
    idx = np.where(sel)[0]
    peaks = [i for i in idx[1:-1] if spec[i] > spec[i - 1] and spec[i] > spec[i + 1] and spec[i] > 0.05 * spec[idx].max()]
    out = []
    for i in peaks:                      # parabolic refinement
        a, b, c = np.log(spec[i - 1]), np.log(spec[i]), np.log(spec[i + 1])
        out.append(f[i] + 0.5 * (a - c) / (a - 2 * b + c) * (f[1] - f[0]))
    return np.array(out)


def _cavity(n, a, steps, regions=None, comp=2):
    s = OracleSim((n, n, n), a, pml=0.0, nsets=1)
    if regions is not None:
        s.set_regions(*regions, [np.ones((n + 1, n + 1, n + 1), dtype=np.uint8)] * 3)
    L = n / a
    # small off-centre volume source, broadband pulse (integrated source = polarisation kick)
    s.add_gaussian_source(comp, [0.21 * L, 0.33 * L, 0.27 * L], [0.34 * L, 0.46 * L, 0.4 * L], 1.0, 0.9, 0.35, 0.0, 0.0, 4.2, True)
    s.add_monitors([[0.37 * L, 0.29 * L, 0.61 * L]], comp)
    s.run(steps, 1)
    return s.monitors()[:, 0, 0], s.dt


def _yee_mode(l, m, n_, L, a, dt):
    dx = 1.0 / a
    rhs = sum((np.sin(np.pi * q / L * dx / 2) / dx) ** 2 for q in (l, m, n_))
    return 2 / dt * np.arcsin(dt * np.sqrt(rhs)) / (2 * np.pi)


def test_vacuum_cavity_modes_follow_yee_dispersion():
    n, a = 16, 16.0
    sig, dt = _cavity(n, a, 6000)
    got = _spectrum_peaks(sig[400:], dt, 0.5, 1.25)
    # Ez modes of the unit PEC cube: l,m >= 1, n >= 0
    want = sorted({round(_yee_mode(l, m, q, 1.0, a, dt), 6) for l in (1, 2) for m in (1, 2) for q in (0, 1, 2)})
    want = [f for f in want if 0.5 < f < 1.25]
    assert len(got) >= 3
    for g in got:
        assert min(abs(g - w) for w in want) < 2e-3, (g, want)
    # the lowest mode must sit at the DISCRETE frequency, measurably below the continuum sqrt(2)/2
    f110 = _yee_mode(1, 1, 0, 1.0, a, dt)
    assert abs(got[0] - f110) < 2e-5 and abs(got[0] - np.sqrt(2) / 2) > 4e-4


@pytest.mark.parametrize("pole", [(1.4, 0.0, 1.1, 0), (1e-10, 0.0, 0.9e20, 1)])
def test_dispersive_cavity_resonances(pole):
    """eps(f) f^2 = f_vac^2 for a cavity filled with one Lorentz (or Drude) pole, eps_inf = 2.25."""
    n, a = 16, 16.0
    eps_inf = 2.25
    f0, gam, sg, drude = pole
    sig, dt = _cavity(n, a, 8000, regions=(eps_inf, [eps_inf], [[pole]]))
    got = _spectrum_peaks(sig[400:], dt, 0.2, 1.2)
    fv = _yee_mode(1, 1, 0, 1.0, a, dt)
    if drude:
        fp2 = sg * f0 ** 2
        want = [np.sqrt((fv ** 2 + fp2) / eps_inf)]
    else:
        # eps_inf f^4 - (eps_inf f0^2 + sigma f0^2 + fv^2) f^2 + fv^2 f0^2 = 0
        b = eps_inf * f0 ** 2 + sg * f0 ** 2 + fv ** 2
        disc = np.sqrt(b * b - 4 * eps_inf * fv ** 2 * f0 ** 2)
        want = [np.sqrt((b - disc) / (2 * eps_inf)), np.sqrt((b + disc) / (2 * eps_inf))]
    for w in want:
        if 0.2 < w < 1.2:
            assert min(abs(got - w)) < 0.012 * w, (w, got)


def test_pml_absorbs_and_lossy_medium_is_passive():
    n, a = 24, 6.0
    L = n / a
    s = OracleSim((n, n, n), a, pml=1.0, nsets=1)
    mask = np.zeros((n + 1, n + 1, n + 1), dtype=np.uint8)
    mask[14:, :, :] = 1
    s.set_regions(1.0, [2.0], [[(1.2, 0.3, 0.8, 0)]], [mask] * 3)
    s.add_gaussian_source(0, [0, 0, 1.2], [L, L, 1.2], 1.0, 0.5, 1.0, 0.0, 0.5, 8.5, True)
    s.add_monitors([[L / 2, L / 2, L / 2]], 0)
    energy = []
    for i in range(420):
        s.step()
        if i % 20 == 0:
            energy.append(sum(float((s.field(k, c) ** 2).sum()) for k in ("E", "H") for c in range(3)))
    assert np.isfinite(energy).all()
    tail = energy[7:]                      # the source is off after step ~102
    assert all(b <= a_ * 1.0000001 for a_, b in zip(tail, tail[1:])), energy
    assert energy[-1] < 1e-4 * max(energy)


def test_cw_waveform_and_steady_state():
    """meep::continuous_src_time as the reference uses it for CW_source (disp.cpp:615-619): exp(-i w t)/(-i w) with
    tanh turn-on/off; zero outside [start, end]; a vacuum run settles into a sinusoid at the drive frequency."""
    n, a = (10, 10, 40), 8.0
    f, width, t0, t1 = 0.45, 1.0, 0.5, 30.0
    o = OracleSim(n, a, pml=1.0, nsets=2)
    o.add_cw_source(0, [0, 0, 1.5], [n[0] / a, n[1] / a, 1.5], 1.0, f, width, t0, t1, 3.0)
    assert o.last_source_time() == t1
    w = 2 * np.pi * f
    for t in (0.0, 0.49, 0.5, 3.0, 15.0, 29.0, 30.01, 40.0):
        ts, te = (t - t0) / width - 3.0, (t1 - t) / width - 3.0
        want = 0.0 if (t < t0 or t > t1) else np.exp(-1j * w * t) / (-1j * w) * (1 + np.tanh(ts)) * (1 + np.tanh(te)) / 4
        got = o.dipole(0, t)
        assert abs(got - want) <= 1e-15 * max(1.0, abs(want)), t
    o.add_monitors([[n[0] / a / 2, n[1] / a / 2, 3.0]], 0)
    steps = int(25.0 / o.dt)
    o.run(steps, 1)
    m = o.monitors()[:, 0, 0] + 1j * o.monitors()[:, 0, 1]
    tail = m[int(12.0 / o.dt):]                     # ramp finished (t0 + ~6 widths), well before turn-off
    assert np.abs(tail).min() > 0.9 * np.abs(tail).max() > 0.0            # complex envelope is flat: a pure tone
    phase = np.unwrap(np.angle(tail))
    slope = np.polyfit(np.arange(len(tail)) * o.dt, phase, 1)[0]
    assert abs(-slope / (2 * np.pi) - f) < 2e-3
