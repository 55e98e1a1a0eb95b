"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, mean, max, share of the total."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = {h: k for k, h in enumerate(rows[hi])}
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= H["Metric Value"] or r[H["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[H["Metric Value"]].replace(",", ""))
    unit = r[H["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    acc.setdefault(r[H["Kernel Name"]], []).append(v)
tot = sum(sum(v) for v in acc.values())
print(" ".join(sys.argv[2:]))
print("(per-launch device time, cold-cache and serialised: compare shares)")
for name, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name[:96]:96s} n={len(v):4d} mean={sum(v)/len(v):9.1f}us max={max(v):9.1f}us share={100*sum(v)/tot:5.1f}%")
