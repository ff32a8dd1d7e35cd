#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file list.csv ...).
usage: ncu_launch_shares.py list.csv ["# header line" ...]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    k = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    v = float(r[hdr.index("Metric Value")].replace(",", ""))
    u = r[hdr.index("Metric Unit")]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(t for _, t in agg.values())
for h in sys.argv[2:]:
    print(h)
print("%-45s %8s %12s %7s %10s" % ("kernel", "launches", "total ms", "share", "avg us"))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-45s %8d %12.3f %6.1f%% %10.1f" % (k[:45], n, t, 100 * t / tot, 1e3 * t / n))
