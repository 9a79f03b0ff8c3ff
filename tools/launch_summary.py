"""Group an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel and grid:
python tools/launch_summary.py launches.csv > profiles/<name>_summary.txt"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) != len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = (r[ix["Kernel Name"]][:64], r[ix["Grid Size"]], r[ix["Block Size"]])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
for (name, grid, block), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:64s} grid={grid:>14s} block={block:>12s} launches={n:3d} total_ms={ms:9.3f} avg_ms={ms / n:8.3f} share={100 * ms / tot:5.1f}%")
