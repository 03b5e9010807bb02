"""Dev tool: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, idc = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv or int(r[idc]) < skip:
        continue
    name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:70]
    agg.setdefault(name, []).append(float(r[mv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'mean us':>9s} {'total ms':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} {len(v):5d} {sum(v)/len(v)/1e3:9.1f} {sum(v)/1e6:9.3f} {sum(v)/tot:6.1%}")
