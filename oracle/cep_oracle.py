"""TEST INFRASTRUCTURE (only tests/ may import this): NumPy restatement of the front half of the reference's pulse
post-processing class `signal` (/root/reference/scripts/phases.py) -- the part sj_extract_cep computes on the device.
Pinned against the reference class itself: scripts/make_cep_golden.py imports phases.py in the build container (h5py and
matplotlib stubbed), runs `signal` on the committed series and stores its attributes in tests/golden/cep_signal.npz;
tests/test_cep_oracle.py compares this restatement with them."""
import numpy as np

CUT_ALPHA = 0.1          # phases.py:25


def find_peaks(x):
    """scipy.signal.find_peaks without conditions: strict rise, plateau midpoints, strict fall"""
    pk = []
    i, n = 1, len(x)
    while i < n - 1:
        if x[i - 1] < x[i]:
            j = i
            while j + 1 < n - 1 and x[j + 1] == x[i]:
                j += 1
            if x[j + 1] < x[i]:
                pk.append((i + j) // 2)
            i = j
        i += 1
    return np.array(pk, dtype=int)


def fix_angle(theta):
    """phases.py:49-60"""
    if theta > np.pi:
        theta -= 2 * np.pi * np.floor(theta / (2 * np.pi))
        if theta > np.pi:
            theta -= 2 * np.pi
    elif theta <= -np.pi:
        theta += 2 * np.pi * (np.floor(-theta / (2 * np.pi)) + 1)
        if theta > np.pi:
            theta -= 2 * np.pi
    return theta


def fix_angle_seq(angles):
    """phases.py:62-74"""
    angles = np.array(angles, dtype=float)
    n = len(angles)
    for i in range(1, n):
        dist_raw = abs(angles[i] - angles[i - 1])
        if abs(angles[i] + 2 * np.pi - angles[i - 1]) < dist_raw:
            angles[i:] += 2 * np.pi
        elif abs(angles[i] - 2 * np.pi - angles[i - 1]) < dist_raw:
            angles[i:] -= 2 * np.pi
    return angles


def signal_params(t_pts, v_pts):
    """signal.__init__ up to and including _param_est for a series that needs no low-pass (phases.py:289-348, 126-133):
    dict with t0_ind (the guess, :95-106), f0 / f0_ind ('avg' mode, :228-231), f_min / f_max (:108-124), t0_corr, phi_corr,
    slope, intercept."""
    v = np.asarray(v_pts, dtype=float)
    dt = t_pts[1] - t_pts[0]
    pk = find_peaks(v ** 2)
    tot = np.abs(v[pk]).sum() if len(pk) else 0.0
    t0_ind = int((pk * np.abs(v[pk])).sum() / tot) if tot > 0 else 0
    freqs = np.fft.rfftfreq(len(v), d=dt)
    vf = 2 * np.fft.rfft(np.roll(v, -t0_ind))
    mags = np.abs(vf)
    pk_i = int(np.argmax(mags))
    m = 2 * pk_i
    f0 = np.trapz(mags[:m] * freqs[:m]) / np.trapz(mags[:m])
    f0_ind = int(np.rint(f0 / (freqs[1] - freqs[0])))
    cut = mags[f0_ind] * CUT_ALPHA
    f_min = f_max = f0_ind
    for j in range(min(f0_ind, len(mags) - f0_ind - 1)):
        if mags[f0_ind - j] > cut:
            f_min = f0_ind - j - 1
        if mags[f0_ind + j] > cut:
            f_max = f0_ind + j + 1
    def regress(vf_):
        ang = fix_angle_seq(np.angle(vf_[f_min:f_max]))
        x = freqs[f_min:f_max]
        mx, my = x.mean(), ang.mean()
        sl = ((x - mx) * (ang - my)).sum() / ((x - mx) ** 2).sum()
        return sl, my - sl * mx
    slope, icpt = regress(vf)
    # _param_est recentres once (MAX_PARAM_EVALS = 1) when the fitted delay is large against the band (phases.py:148-155):
    # the series is rolled by the fitted delay as well and the regression repeated on the same band
    if abs(-slope / (2 * np.pi)) > 2 / (f_max - f_min):
        t0_ind2 = t0_ind + int(-slope / (2 * np.pi) / dt)
        slope, icpt = regress(2 * np.fft.rfft(np.roll(v, -t0_ind2)))
    return dict(t0_ind=t0_ind, f0=f0, f0_ind=f0_ind, f_min=f_min, f_max=f_max, slope=slope, intercept=icpt,
                t0_corr=-slope / (2 * np.pi), phi_corr=fix_angle(icpt))
