"""Instruction mix of the per-tile loop bodies in an `ncu --page source --csv` export:
python tools/ncu_src_mix.py file.src.csv <executions per instruction of the loop, e.g. tiles x warps> [tolerance]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
base = float(sys.argv[2])
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
def f(r, k):
    try:
        return float(r[ix[k]])
    except ValueError:
        return 0.0
sel = [r for r in body if f(r, "Instructions Executed") >= base * (1 - tol)]
tot = sum(f(r, "Instructions Executed") for r in sel) / base
print(f"{len(sel)} SASS lines at >= {base:.0f} executions; {tot:.1f} instructions per loop pass; samples {sum(f(r, '# Samples') for r in sel):.0f}")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(f(r, k) for r in sel) for k in stalls}
print("stalls:", ", ".join(f"{k[6:]}={int(v)}" for k, v in sorted(agg.items(), key=lambda t: -t[1])[:9]))
c = collections.Counter()
for r in sel:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']].strip())
    c[s.split()[0].split('.')[0]] += f(r, "Instructions Executed") / base
print("mix:", ", ".join(f"{k} {v:.1f}" for k, v in c.most_common(40)))
