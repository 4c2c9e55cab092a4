"""Executed-instruction mix of the first kernel in an ncu report (needs the
report captured with --import-source on): python tools/ncu_opmix.py rep [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
stall = collections.Counter()
total = 0
heavy = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ci["Instructions Executed"]])
        smp = int(r[ci["Warp Stall Sampling (All Samples)"]])
    except ValueError:
        continue
    toks = r[ci["Source"]].split()
    op = toks[0]
    if op.startswith("@"):
        op = toks[1]
    op = op.split(".")[0]
    ops[op] += n
    stall[op] += smp
    total += n
    heavy.append((smp, n, r[ci["Source"]].strip()))
print("total warp instructions", total)
for op, n in ops.most_common(top):
    print(f"{op:10s} {n:14d} {100.0 * n / total:6.2f}%   stall samples {stall[op]}")
print("--- top stalled instructions")
for smp, n, src in sorted(heavy, reverse=True)[:15]:
    print(smp, n, src)
