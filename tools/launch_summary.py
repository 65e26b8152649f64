"""Condenses an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) into per-kernel totals:
  python tools/launch_summary.py gpurun_out/launches.csv "comment line" > profiles/rNN_launches_summary.csv"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(r[iu], 1.0)
    name = re.sub(r"\s+", " ", r[ik]).replace("mo::<", "").replace(",", ";")[:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("kernel,launches,total_us,share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.4f" % (k, a[0], a[1], a[1] / tot))
