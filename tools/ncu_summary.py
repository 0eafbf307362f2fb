"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the raw sequence."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("hmvae::", "").replace("void ", "")
    try:
        seq.append((name, float(r[vi].replace(",", "")) / 1000.0))
    except ValueError:
        pass
agg, cnt = collections.Counter(), collections.Counter()
for n, v in seq:
    n = re.sub(r"<.*", "", n)
    agg[n] += v
    cnt[n] += 1
tot = sum(agg.values())
print("total %.1f us over %d launches" % (tot, len(seq)))
for k, v in agg.most_common(25):
    print("%9.1f us %5.1f%% x%-3d %s" % (v, v / tot * 100, cnt[k], k[:70]))
if len(sys.argv) > 2:
    print("--- sequence")
    for n, v in seq:
        if any(s in n for s in sys.argv[2].split(",")):
            print("%9.1f us  %s" % (v, n[:80]))
