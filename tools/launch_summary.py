"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel (not a pytest file).
usage: python tools/launch_summary.py gpurun_out/X_launches.csv [top]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    n = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("bdetr::", "")[:60]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:62s} {v[0]:5d} {v[1] / 1e3:9.1f} us {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:7.1f}")
print(f"total {tot / 1e3:.1f} us in {sum(v[0] for v in agg.values())} launches")
