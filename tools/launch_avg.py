"""Average kernel durations from an ncu launch-list CSV, optionally only launches with a given grid size.
usage: python tools/launch_avg.py <launches.csv>"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, gs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0][-48:]
    agg.setdefault((name, r[gs]), []).append(float(r[mv].replace(",", "")))
for (name, g), v in agg.items():
    v2 = sorted(v)
    print("%-50s grid=%-16s n=%3d avg=%9.1f med=%9.1f min=%9.1f ns" % (name, g, len(v), sum(v) / len(v), v2[len(v2) // 2], v2[0]))
