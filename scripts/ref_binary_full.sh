#!/bin/bash
# The reference's unmodified sim_geom (host/_ref/sim_geom_meep: main.cpp + disp.cpp over host/meep_compat) and this
# repository's C++ host on the full Au_graphene_box production run (181^3, ~33 k steps), timed, outputs compared.
G="--conf-file scenes/Au_graphene_box/params.conf --geom-file scenes/Au_graphene_box/junc.geom"
mkdir -p /tmp/o1 /tmp/o2 gpurun_out
L=gpurun_out/s3_compat_full.log
nproc > $L
( time timeout 500 host/_ref/sim_geom_meep $G --out-dir /tmp/o1 ) 2>&1 | grep -v "complete$" | grep -v "^	*(" | tail -14 >> $L
( time timeout 500 host/_ref/sim_geom $G --out-dir /tmp/o2 ) 2>&1 | grep -v "complete$" | grep -v "^	*(" | tail -14 >> $L
python - >> $L 2>&1 <<PY
import numpy as np, sys
sys.path.insert(0, ".")
from sim_juncs_b200 import hdf5
a = hdf5.File("/tmp/o1/field_samples.h5"); b = hdf5.File("/tmp/o2/field_samples.h5")
worst = 0.0; nrm = 0.0; n = 0
for pt in a["cluster_0"].keys():
    if not pt.startswith("point_"): continue
    x = a["cluster_0"][pt]["time"].read(); y = b["cluster_0"][pt]["time"].read()
    worst += float(np.sum((x["Re"] - y["Re"]) ** 2 + (x["Im"] - y["Im"]) ** 2)); nrm += float(np.sum(y["Re"] ** 2 + y["Im"] ** 2)); n += 1
print(n, "monitors,", len(x), "samples each; rel L2 reference-binary-on-B200 vs own C++ host:", (worst / nrm) ** 0.5)
PY
cat $L
