"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, and (optional) the ordered list."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[1] for r in rows)
agg = collections.OrderedDict()
for n, us, g, b in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += us
print(f"total {tot/1e3:.2f} ms over {len(rows)} launches")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*us/tot:6.2f}%  n={c:4d}  avg={us/c:9.1f} us  {n[:100]}")
if len(sys.argv) > 2:
    for n, us, g, b in rows:
        print(f"{us:9.1f} us  {g:>14s} {b:>12s}  {n[:80]}")
