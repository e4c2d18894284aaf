"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name -> markdown (profiles/)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
idx = {h: i for i, h in enumerate(rows[hdr])}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    t = tot[r[idx["Kernel Name"]]]
    t[0] += 1
    t[1] += ms
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "# ncu launch list")
print("# per-launch times are cold-cache and serialised: compare SHARES\n")
print(f"total kernel time {total:.1f} ms over {n} launches\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
mine = 0.0
for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| {k[:90]} | {c} | {ms:.2f} | {ms / total:.3f} |")
for k, (c, ms) in tot.items():
    if "mv::" in k or k.startswith("mv::") or "tapgemm" in k or "conv3" in k or "head3" in k or "wgrad" in k or "lpx" in k or "moe_lw" in k \
            or "upsample2x" in k or "avgpool3s2" in k or "head_grad_pack" in k or "colsum" in k or "scale_dact" in k or "poe_" in k:
        mine += ms
print(f"\nkernels of this library: {mine / total:.3f} of the step's kernel time")
