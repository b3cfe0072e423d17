"""Output contract pieces that need no GPU: the reference's quirky `frequency` transform and naming."""
import numpy as np

from sim_juncs_b200.output import cluster_name, point_name, reference_fft


def test_fft_of_sine_matches_reference_kat():
    # src/main_test.cpp:20-34: x[k] = sin(4 pi k / 16): bin 2 = -8i (the real transform there; the complex
    # transform has the same bin for a real input)
    n = 16
    x = np.sin(4 * np.pi * np.arange(n) / n)
    f = reference_fft(x)
    assert len(f) == 16
    assert abs(f[2] - (-8j)) < 1e-12 and abs(f[n - 2] - 8j) < 1e-12
    assert all(abs(f[k]) < 1e-12 for k in range(n) if k not in (2, n - 2))


def test_fft_truncates_to_power_of_two_but_keeps_full_length_phase():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(175) + 1j * rng.standard_normal(175)     # the production run stores 175 saves
    f = reference_fft(x)
    assert len(f) == 128                                             # 1600 x 128 out
    k = 5
    want = sum(x[n] * np.exp(-2j * np.pi * n * k / 175) for n in range(128))
    assert abs(f[k] - want) < 1e-10
    km = -3
    want = sum(x[n] * np.exp(-2j * np.pi * n * km / 175) for n in range(128))
    assert abs(f[128 + km] - want) < 1e-10


def test_power_of_two_equals_numpy_fft():
    x = np.random.default_rng(1).standard_normal(64)
    assert np.allclose(reference_fft(x), np.fft.fft(x), atol=1e-10)


def test_zero_padded_names():
    # disp.cpp:873-917: digits = floor(log10(count)) + 1
    assert point_name(7, 1600) == "point_0007" and point_name(1599, 1600) == "point_1599"
    assert cluster_name(3, 40) == "cluster_03" and cluster_name(40, 40) == "cluster_40"
    assert point_name(1, 2) == "point_1" and cluster_name(0, 1) == "cluster_0"


def test_make_dec_str_reference_kats():
    # src/main_test.cpp:5-15
    from sim_juncs_b200.output import field_dump_name, make_dec_str
    assert make_dec_str(0.25, 1, 2, ".") == "0.25"
    assert make_dec_str(0.25, 2, 2, "_") == "00_25"
    assert make_dec_str(123.5, 2, 1) is None                      # more integer digits than allotted: the reference's -2
    # disp.cpp:709-713: digits from ttot and dt; run.conf-like numbers (ttot = 23.4, dt = 0.1)
    assert field_dump_name(1.5, 23.4, 0.1) == "ex-01.50.h5"
    assert field_dump_name(0.0, 174.9, 0.04972) == "ex-000.000.h5"


def test_centred_interpolation_of_yee_arrays():
    from sim_juncs_b200.output import centred
    n = 5
    k, j, i = np.meshgrid(np.arange(n + 1.0), np.arange(n + 1.0), np.arange(n + 1.0), indexing="ij")
    # a field linear in the coordinates is reproduced exactly at the pixel centres by the four-point mean
    for comp, off in ((0, (0.5, 0.0, 0.0)), (1, (0.0, 0.5, 0.0)), (2, (0.0, 0.0, 0.5))):
        f = 2.0 * (i + off[0]) - 3.0 * (j + off[1]) + 0.5 * (k + off[2])      # value at the component's own Yee point
        c = centred(f, comp)
        assert c.shape == (n, n, n)
        x, y, z = np.meshgrid(np.arange(n) + 0.5, np.arange(n) + 0.5, np.arange(n) + 0.5, indexing="ij")
        assert np.allclose(c, 2.0 * x - 3.0 * y + 0.5 * z, atol=1e-12)
