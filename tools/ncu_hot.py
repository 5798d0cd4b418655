"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` output (with -lineinfo source mapping)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = None
for i, r in enumerate(rows):
    if "Source" in r and any("Samp" in c for c in r):
        h = i
        break
hdr = rows[h]
isrc = hdr.index("Source")
isamp = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"][0]
iaddr = hdr.index("Address") if "Address" in hdr else 0
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_")]
data = []
for r in rows[h + 1:]:
    if len(r) <= isamp:
        continue
    try:
        s = float(r[isamp] or 0)
    except ValueError:
        continue
    data.append((s, r))
tot = sum(d[0] for d in data) or 1
print("columns:", [c for c in hdr if c][:40])
print("total samples", tot)
for s, r in sorted(data, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    top = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2] if stall_cols else []
    print(f"{100*s/tot:5.1f}%  {r[isrc][:90]:90s} {top}")
