"""Static size of the loops of one kernel in the built library: python scripts/sass_loops.py <mangled-substring>"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "sim_juncs_b200", "lib", "libsimjuncs_b200.so")],
                     capture_output=True, text=True).stdout
cur, ins = None, {}
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); ins[cur] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m and cur:
        ins[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, L in ins.items():
    if sys.argv[1] not in name:
        continue
    print(name, len(L), "instructions;", sum("LDG" in t for _, t in L), "LDG", sum("STG" in t for _, t in L), "STG",
          sum("LDL" in t or "STL" in t for _, t in L), "local")
    for a, t in L:
        m = re.search(r"BRA\s+(?:U,!UPT,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            b = int(m.group(1), 16)
            body = [x for x in L if b <= x[0] <= a]
            print("   loop %5x..%5x %4d instr  LDG %d STG %d local %d" % (b, a, len(body), sum("LDG" in t for _, t in body),
                  sum("STG" in t for _, t in body), sum("LDL" in t or "STL" in t for _, t in body)))
