"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file F) by kernel: python scripts/launch_summary.py F [skip_first_n]"""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if skip < 0:            # -k: keep only the last 1/k of the launches (k identical steps were captured)
    skip = (len(rows) - 1) * (-skip - 1) // (-skip)
agg = {}
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki])[:90]
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("%d launches, %.1f us total" % (sum(a[0] for a in agg.values()), tot))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%9.1f us %5.1f%% %5d x %7.2f us  %s" % (t, 100 * t / tot, c, t / c, name))
