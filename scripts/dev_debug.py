import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import OracleSim
from sim_juncs_b200 import Sim
n=(20,20,20); a=5.0
for steps,mon in ((60,0),(210,0),(30,1),(210,1)):
    o = OracleSim(n, a, pml=1.0, nsets=2); g = Sim(n, a, pml=1.0, n_sets=2)
    args=(1,[0,0,1],[4,4,1],1.0,0.4,1.5,0.3,1.0,19.0,True)
    o.add_gaussian_source(*args); g.add_gaussian_source(*args)
    if mon:
        mm=[[2,2,2],[4/3,4,4/3],[0.01,0.02,0.03]]
        o.add_monitors(mm,1); g.add_monitors(mm,1)
    o.run(steps,1); g.run(steps,1)
    if mon:
        mo,mg=o.monitors(),g.monitors()
        print("  mon diff per monitor", np.abs(mo-mg).max(axis=(0,2)), "max", np.abs(mo).max(axis=(0,2)))
    print("steps",steps,"mon",mon)
    for q in range(2):
        for c in range(3):
            for nm,kind,cc in (("E","E",c),("H","H",3+c)):
                A=g.field(cc,q); B=o.field(kind,c,q)
                d=np.abs(A-B); 
                if d.max()>1e-14:
                    w=np.unravel_index(d.argmax(), d.shape)
                    print("  set",q,nm,c,"max diff %.3e at kji"%d.max(),w,"gpu",A[w],"orc",B[w], "n_bad", (d>1e-13).sum())
