"""Per-frame kernel breakdown from an ncu launch list (gpurun_out/launches.csv): one streaming frame = the launches
between two consecutive conv_cin1 kernels."""
import csv, sys
from collections import defaultdict
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
names = [r[4].split('(')[0].replace('rstnet::', '').replace('void ', '') for r in rows]
starts = [i for i, n in enumerate(names) if 'conv_cin1' in n]
a, b = starts[0], starts[1]
agg = defaultdict(lambda: [0, 0.0]); tot = 0.0
for r, n in zip(rows[a:b], names[a:b]):
    t = float(r[14]) / 1000
    agg[n][0] += 1; agg[n][1] += t; tot += t
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:60]:60s} {c:4d} {t:9.1f} us")
print(f"total {tot:.1f} us over {b - a} launches")
if "-v" in sys.argv:
    for r, n in zip(rows[a:b], names[a:b]):
        if 'gemm' in n:
            print(f"{n[:40]:40s} grid={r[8]:18s} {float(r[14]) / 1000:8.1f}")
