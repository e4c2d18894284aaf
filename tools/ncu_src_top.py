"""Hottest SASS instructions of an `ncu --page source --csv` export: python tools/ncu_src_top.py file.src.csv [n]
Prints total executed warp-instructions, then the top-n instructions by executed count and by stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
def f(r, k):
    try:
        return float(r[ix[k]])
    except ValueError:
        return 0.0
tot = sum(f(r, "Instructions Executed") for r in body)
smp = sum(f(r, "# Samples") for r in body)
print(f"instructions executed {tot:.0f}, samples {smp:.0f}, SASS lines {len(body)}")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
print("-- by executed count")
for r in sorted(body, key=lambda r: -f(r, "Instructions Executed"))[:n]:
    print(f"{int(f(r, 'Instructions Executed')):10d} {int(f(r, '# Samples')):7d}  {r[ix['Address']][-6:]}  {r[ix['Source']][:90]}")
print("-- by samples (dominant stalls)")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:n]:
    st = sorted(((f(r, k), k) for k in stalls), reverse=True)[:2]
    print(f"{int(f(r, '# Samples')):7d} {int(f(r, 'Instructions Executed')):10d}  {r[ix['Address']][-6:]}  {r[ix['Source']][:70]:70s} {st[0][1]}={int(st[0][0])} {st[1][1]}={int(st[1][0])}")
