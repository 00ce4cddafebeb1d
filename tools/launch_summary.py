"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/launch_summary.py gpurun_out/launches.csv [--seq KERNEL_SUBSTR]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        rows.append((row["Kernel Name"], us))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    for name, us in rows:
        k = name.split("(")[0][-48:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:50s} n={n:5d} total={t / 1e3:9.3f} ms mean={t / n:9.1f} us {100 * t / tot:5.1f}%")
    if "--seq" in sys.argv:
        sub = sys.argv[sys.argv.index("--seq") + 1]
        print("per-launch us of", sub, ":", " ".join(f"{us:.0f}" for n, us in rows if sub in n))


if __name__ == "__main__":
    main()
