"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one step.
usage: python scripts/summarize_launches.py launches.csv launches_per_step > summary.md"""
import collections
import csv
import re
import sys

path, per = sys.argv[1], int(sys.argv[2])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))[-per:]


def us(row):
    t = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)


groups = collections.OrderedDict()
tot = 0.0
for row in rows:
    n = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void evc::", "").replace("<unnamed>::", "")
    groups.setdefault((n, row["Grid Size"]), []).append(us(row))
    tot += us(row)
print(f"# last {per} launches of {path} (one step), serialized cold-cache ncu times: compare shares\n")
print(f"total kernel time {tot/1e3:.3f} ms\n")
print("| kernel | grid | launches | avg us | total us | share |")
print("|---|---|---:|---:|---:|---:|")
for (n, g), v in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{n}` | {g} | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |")
