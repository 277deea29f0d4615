"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals (and, with -v, the sequence)."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("dbev::<unnamed>::", "dbev::")
        seq.append((name[:70], us))
    return seq


def main():
    seq = load(sys.argv[1])
    agg = collections.OrderedDict()
    for k, t in seq:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(t for _, t in seq)
    print("total %.1f us, %d launches" % (tot, len(seq)))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %4d %9.1f us %5.1f%%" % (k, n, t, 100 * t / tot))
    if "-v" in sys.argv:
        for i, (k, t) in enumerate(seq):
            print(i, k, round(t, 1))


if __name__ == "__main__":
    main()
