"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [skip]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr, seq = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            t = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if d["Metric Unit"] == "ns":
            t /= 1e3
        elif d["Metric Unit"] == "ms":
            t *= 1e3
        seq.append((re.sub(r"\(.*", "", d["Kernel Name"])[:70], t))
seq = seq[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in seq:
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(t for _, t in seq)
print("launches %d, total %.1f us" % (len(seq), tot))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.1f us %5.1f %% %5d x  %s" % (t, 100 * t / tot, c, n))
