"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, time, share.

    python tools/launch_summary.py launches.csv [top] [--steps K] [--out trimmed.csv]

--steps K keeps only the last K WHOLE train steps of the capture (a step ends with its hn::adam_flat_kernel launch), so
that the shares are those of complete steps whatever window ncu captured; --out writes those launches back as CSV."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines)), lines[0]


def micros(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1000 if u == "ns" else v * 1000 if u == "ms" else v


def whole_steps(rows, k):
    ends = [i for i, r in enumerate(rows) if "adam_flat_kernel" in r["Kernel Name"]]
    if len(ends) < k + 1:
        raise SystemExit("only %d optimiser launches in the capture: cannot cut %d whole steps" % (len(ends), k))
    return rows[ends[-k - 1] + 1: ends[-1] + 1]


def main(argv):
    path = argv[0]
    top = int(argv[1]) if len(argv) > 1 and not argv[1].startswith("--") else 24
    steps = int(argv[argv.index("--steps") + 1]) if "--steps" in argv else 0
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    rows, header = load(path)
    if steps:
        rows = whole_steps(rows, steps)
    if out:
        with open(out, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()), quoting=csv.QUOTE_ALL)
            w.writeheader()
            w.writerows(rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        k = row["Kernel Name"][:72]
        agg[k][0] += 1
        agg[k][1] += micros(row)
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
    print("\ntotal %.1f us over %d launches%s" % (tot, sum(v[0] for v in agg.values()),
                                                 " (%d whole steps)" % steps if steps else ""))


if __name__ == "__main__":
    main(sys.argv[1:])
