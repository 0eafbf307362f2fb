"""Summarises the launch list of tools/conv_layers.py: for each (kernel, grid) in order of first appearance, the MIN duration
(us) over its launches -- i.e. the warm time.  Usage: python tools/ncu_layers.py launches.csv"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
order, best, cnt = [], {}, {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("hmvae::", "").replace("void ", "")
    name = re.sub(r"<.*", "", name)
    if name.startswith("at::") or "elementwise" in name or "distribution" in name:
        continue
    key = (name, r[gi] if gi is not None else "")
    try:
        v = float(r[vi].replace(",", "")) / 1000.0
    except ValueError:
        continue
    if key not in best:
        order.append(key)
        best[key], cnt[key] = v, 0
    best[key] = min(best[key], v)
    cnt[key] += 1
tot = {}
for key in order:
    print("%8.1f us  x%-3d %-28s %s" % (best[key], cnt[key], key[0][:28], key[1]))
    tot[key[0]] = tot.get(key[0], 0.0) + best[key]
print("--- sum of warm times per kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%8.1f us  %s" % (v, k))
