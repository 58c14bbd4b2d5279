"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the run)."""
import collections
import csv
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"]
        k = k[:k.index("(")] if "(" in k else k
        a = agg.setdefault(k[:80], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("%-82s %5s %12s %6s %10s" % ("kernel", "n", "total_us", "share", "avg_us"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-82s %5d %12.1f %5.1f%% %10.1f" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0] / 1e3))
    print("total %.1f us over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
