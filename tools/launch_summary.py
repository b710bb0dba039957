"""Summarise an ncu --metrics gpu__time_duration.sum CSV launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
rows = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("unnamed>::", "")
    tot[name] += v
    cnt[name] += 1
    rows.append((name, v, row.get("Grid Size", "")))
T = sum(tot.values())
print(f"total {T:.1f} us over {len(rows)} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:3d}  {k}")
if len(sys.argv) > 2:
    for n, v, g in rows:
        if sys.argv[2] in n:
            print(f"{v:9.1f} us  grid {g}  {n}")
