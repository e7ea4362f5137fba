"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^.*::", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        seq.append((name, v, row["Grid Size"], row["Block Size"]))
    seq = seq[skip:]
    for name, v, _, _ in seq:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':45s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{n[:45]:45s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {a[1] / tot * 100:5.1f}%")
    print(f"{'total':45s} {len(seq):8d} {tot:10.1f}")
    return seq


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
