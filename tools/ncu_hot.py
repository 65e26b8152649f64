"""Top SASS instructions of an .ncu-rep by stall samples (needs --import-source on):
  python tools/ncu_hot.py gpurun_out/prof.ncu-rep [N] [kernel-substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
data = []
kern = ""
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        d["_k"] = kern
        d["_i"] = len(data)
        data.append(d)
if len(sys.argv) > 3:
    data = [d for d in data if sys.argv[3] in d["_k"]]
tot = sum(int(d["# Samples"] or 0) for d in data) or 1
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(d[h] or 0) for d in data) for h in stall_cols}
print("stall mix:", ", ".join("%s=%.1f%%" % (h[6:], 100.0 * v / tot) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for d in sorted(data, key=lambda d: -int(d["# Samples"] or 0))[:top]:
    st = sorted(((h[6:], int(d[h] or 0)) for h in stall_cols), key=lambda x: -x[1])[:2]
    print("%5d %5.1f%% #%-5d %-60s %s  exec=%s shconf=%s" % (int(d["# Samples"]), 100.0 * int(d["# Samples"]) / tot, d["_i"], d["Source"].strip()[:60],
                                                     ",".join("%s:%d" % s for s in st), d["Instructions Executed"], d.get("L1 Wavefronts Shared Excessive", "")))
