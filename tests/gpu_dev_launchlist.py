"""Dev helper: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]; kn = H.index("Kernel Name"); mv = H.index("Metric Value"); mu = H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", ""))
    if r[mu] == "ns": v /= 1000.0
    elif r[mu] == "ms": v *= 1000.0
    a = agg.setdefault(r[kn][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("total kernel time %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-70s n=%6d total %10.1f us  avg %8.2f us  %5.1f%%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
