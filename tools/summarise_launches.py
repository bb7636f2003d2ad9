"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: python tools/summarise_launches.py gpurun_out/rXX_launches.csv "header comment" > profiles/rXX_launches_summary.csv"""
import collections
import csv
import re
import sys

path, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
    name = re.sub(r"\(.*", "", r[ik]).strip()
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
if note:
    print("# " + note)
print(f"# total {total / 1e6:.2f} ms over {n} launches (per-launch times are cold-cache and serialised: compare SHARES with bench.py's `kernels`)")
print("kernel,launches,total_ms,share,avg_us")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{c},{t / 1e6:.3f},{t / total:.4f},{t / c / 1e3:.1f}")
