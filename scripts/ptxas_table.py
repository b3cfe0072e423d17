"""Registers / spills per kernel from the ptxas log of the last build: python scripts/ptxas_table.py [filter]"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "sim_juncs_b200", "_build", "sj_engine.ptxas.log")).read()
ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Used (\d+) registers", log)
spill = dict((a, (int(b), int(c), int(d))) for a, b, c, d in re.findall(
    r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log))
dem = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.split("\n")
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for (n, r), d in zip(ents, dem):
    d = re.sub(r"\(.*", "", d)
    if flt in d:
        print("%-62s regs %3s  stack/spill-st/spill-ld %s" % (d[:62], r, spill.get(n)))
