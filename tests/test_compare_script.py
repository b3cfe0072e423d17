"""scripts/compare_with_meep.py on two files written by the package's own encoder: identical files pass, a perturbed one
fails with the monitor named."""
import os
import subprocess
import sys

import numpy as np


def _write(path, scale):
    from sim_juncs_b200.output import write_field_samples_h5
    rng = np.random.default_rng(3)
    n = 64
    d = {"info/time_bounds": np.array([0.0, 10.0, 0.1]), "info/n_clusters": np.array([1], dtype=np.uint64),
         "info/n_time_points": np.array([n], dtype=np.uint64), "cluster_0/locations": np.array([[1.0, 2.0, 3.0], [2.0, 2.0, 3.0]])}
    for i in range(2):
        s = rng.standard_normal((n, 2))
        if i == 1:
            s = s * scale
        d["cluster_0/point_%d/time" % i] = s
        d["cluster_0/point_%d/frequency" % i] = s[:32]
    write_field_samples_h5(d, path)


def test_compare_script_pass_and_fail(root, tmp_path):
    a, b, c = (str(tmp_path / n) for n in ("a.h5", "b.h5", "c.h5"))
    _write(a, 1.0)
    _write(b, 1.0)
    _write(c, 1.0 + 1e-6)
    exe = [sys.executable, os.path.join(root, "scripts", "compare_with_meep.py")]
    ok = subprocess.run(exe + [a, b], capture_output=True, text=True)
    assert ok.returncode == 0 and "PASS" in ok.stdout
    bad = subprocess.run(exe + [a, c], capture_output=True, text=True)
    assert bad.returncode == 1 and "point_1" in bad.stdout and "above tolerance" in bad.stdout
    loose = subprocess.run(exe + [a, c, "--tol", "1e-4"], capture_output=True, text=True)
    assert loose.returncode == 0
