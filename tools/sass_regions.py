"""Aggregate an `ncu --page source --csv --print-source sass` export into contiguous regions of equal execution count.
usage: python tools/sass_regions.py <source.csv> <kernel-substring> <units (e.g. number of warps' work items)> [min total]"""
import collections
import csv
import sys

path, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
mintot = float(sys.argv[4]) if len(sys.argv) > 4 else 30
rows = list(csv.reader(open(path)))
hd = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
# the kernel name is not in the csv: sections come in launch order, duplicated per view; pick by index given as "#k"
k = int(kern.lstrip("#"))
a, b = hd[k], hd[k + 1]
H = rows[a]
iS, iI, iA = H.index("# Samples"), H.index("Instructions Executed"), H.index("Avg. Threads Executed")
sec = [r for r in rows[a + 1:b] if len(r) > iA and r[iI].isdigit()]
reg = []
for n, r in enumerate(sec):
    c = int(r[iI]) / units
    op = r[1].split()[1] if r[1].startswith("@") else r[1].split()[0]
    smp = int(r[iS] or 0)
    if reg and abs(reg[-1]["c"] - c) < 0.02 * max(1, c):
        reg[-1]["n"] += 1; reg[-1]["s"] += smp; reg[-1]["ops"].append(op); reg[-1]["end"] = n
    else:
        reg.append({"start": n, "end": n, "c": c, "n": 1, "s": smp, "ops": [op], "thr": r[iA]})
tot = sum(r["c"] * r["n"] for r in reg)
stot = sum(r["s"] for r in reg)
print(f"instructions per unit {tot:.0f}; samples {stot}")
for r in reg:
    if r["c"] * r["n"] > mintot:
        cnt = collections.Counter(x.split(".")[0] for x in r["ops"])
        print(f"[{r['start']}-{r['end']}] exec/unit={r['c']:.1f} n={r['n']} total={r['c']*r['n']:.0f} samples={100*r['s']/stot:.1f}% thr={r['thr']} {dict(cnt.most_common(9))}")
