"""SASS evidence of the TMA-staged step kernels: python scripts/sass_excerpt.py > profiles/sass_tma_r2b.txt
(cuobjdump on the built object; no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "sim_juncs_b200", "_build", "sj_tma_f64.o")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
KEY = re.compile(r"\b(UTMALDG[.\w]*|UBLKCP[.\w]*|SYNCS[.\w]*|LDS[.\w]*|STG[.\w]*|LDG[.\w]*|ATOMG[.\w]*|MEMBAR[.\w]*|FENCE[.\w]*|DADD|DMUL|DFMA|DSETP[.\w]*|BAR[.\w]*|SHFL[.\w]*)")
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if "step_tma<double, 224, 16>" not in dem:
        continue
    lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
    ins = [re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-fx]+\*/", "", l)).strip(" ;") for l in lines]
    ins = [i for i in ins if i and not i.startswith("/*")]
    hist = collections.Counter(m.group(1) for i in ins for m in [KEY.search(i)] if m)
    print("=" * 110)
    print(dem.split("(")[0], "-- %d SASS instructions" % len(ins))
    print("  " + "  ".join("%s x%d" % kv for kv in hist.most_common(24)))
    # producer: the TMA issue sequence of one plane load (arm the mbarrier with the byte count, then the box copies)
    for n, i in enumerate(ins):
        if "SYNCS.ARRIVE.TRANS64" in i and any("UTMALDG" in j for j in ins[n:n + 40]):
            print("\n  -- producer, one plane load: mbarrier.arrive.expect_tx, then cp.async.bulk.tensor.3d per array --")
            for j in ins[max(0, n - 2):n + 46]:
                if KEY.search(j) or "R2UR" in j or "UMOV" in j:
                    print("    " + j)
            break
    # consumer: from the mbarrier wait of a plane to its 128-bit stores -- the tightest such window (the interior tiles)
    best = None
    for n, i in enumerate(ins):
        if "SYNCS.PHASECHK" in i:
            win = ins[n:n + 700]
            st = [k for k, j in enumerate(win) if "STG.E.128" in j]
            if len(st) >= 3 and sum("LDS.128" in j for j in win[:st[2]]) >= 8 and not any("SYNCS.PHASECHK" in j for j in win[1:st[2]]):
                if best is None or st[2] < best[1]:
                    best = (n, st[2])
    if best:
        n, m = best
        print("\n  -- consumer, one plane of a tile: mbarrier.try_wait, 128-bit shared loads of the staged boxes, IEEE arithmetic "
              "without contraction, 128-bit stores (%d instructions) --" % (m + 1))
        for j in ins[n:n + m + 1]:
            print("    " + j)
