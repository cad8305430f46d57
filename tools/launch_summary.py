"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py <csv> [title]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("mcl::", "")
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += float(r[14]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[13], 1e-3)
s = sum(v[1] for v in tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# cold-cache, serialised durations under ncu: compare SHARES, not absolutes. %d launches, %.1f us in total." % (len(rows), s))
print("%-40s %6s %12s %7s" % ("kernel", "count", "total us", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %6d %12.1f %6.1f%%" % (k[:40], v[0], v[1], 100 * v[1] / s))
