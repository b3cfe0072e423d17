"""Late-time behaviour: BoundGeom (CUDA) on scenes/tests/graphene_long.conf (res 5, ~5960 steps) against the series the
reference driver produced over the CPU oracle (gpurun_out/reflong_series.npy or tests/golden/ref_graphene_long.npz)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sim_juncs_b200.bound_geom import BoundGeom  # noqa: E402
from sim_juncs_b200.settings import settings_from  # noqa: E402

os.chdir(ROOT)
conf = sys.argv[1] if len(sys.argv) > 1 else "scenes/tests/graphene_long.conf"
extra = sys.argv[2:]
st = settings_from(conf, extra)
bg = BoundGeom(st, None)
bg.run()
got = np.stack(bg.get_field_times(), axis=1)
m = np.abs(got).max(axis=1)
print("steps", bg.n_t_pts, "saves", got.shape)
for i in range(0, len(m), max(len(m) // 12, 1)):
    print(i, "%.3e" % m[i])
np.save("gpurun_out/long_gpu_series.npy", got)
ref_path = "gpurun_out/reflong_series.npy"
if os.path.exists(ref_path) and not extra and conf.endswith("graphene_long.conf"):
    ref = np.load(ref_path)
    n = min(len(ref), len(got))
    for lo in range(0, n, 100):
        hi = min(lo + 100, n)
        print("saves %d-%d rel L2 vs oracle %.3e" % (lo, hi, np.linalg.norm(got[lo:hi] - ref[lo:hi]) / np.linalg.norm(ref[lo:hi])))
