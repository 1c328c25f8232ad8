"""ncu --csv launch list -> per-kernel table (count, total, avg, share).  usage: python profiles/summarize.py file.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    key = (re.sub(r"\(.*", "", row["Kernel Name"])[:44], row.get("Grid Size", ""))
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("%-46s %-16s %6s %12s %10s %7s" % ("kernel", "grid", "n", "total us", "avg us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-46s %-16s %6d %12.1f %10.2f %6.1f%%" % (k[0], k[1], n, t, t / n, 100 * t / tot))
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
