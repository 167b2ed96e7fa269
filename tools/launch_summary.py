"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, time, share."""
import collections
import csv
import sys


def main(path, top=24):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        k = row["Kernel Name"][:72]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
    print("\ntotal %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
