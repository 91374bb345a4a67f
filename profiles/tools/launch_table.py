"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0][-44:]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v for _, v in agg.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("| %s | %d | %.3f | %.1f%% |" % (n, c, v / 1e6, 100.0 * v / tot))
