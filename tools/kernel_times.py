"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv):
mean duration and share per kernel.  python tools/kernel_times.py launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
acc = collections.defaultdict(list)
for r in rows[1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    acc[r[ci["Kernel Name"]]].append(v * scale)
total = sum(sum(v) for v in acc.values())
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:28s} n={len(v):4d} mean={sum(v)/len(v):8.3f} ms share={100*sum(v)/total:5.1f}%")
