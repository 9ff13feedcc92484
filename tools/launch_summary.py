"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import statistics
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        agg[row["Kernel Name"][:70]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-72s %6s %12s %12s %7s" % ("kernel", "n", "median us", "total us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s %6d %12.2f %12.1f %6.1f%%" % (k, len(v), statistics.median(v), sum(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
