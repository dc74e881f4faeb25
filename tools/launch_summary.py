"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the slowest launches."""
import collections
import csv
import re
import sys


def main(path, top=12, thr=None):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg, tot, seq = collections.defaultdict(lambda: [0, 0.0]), 0.0, []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        agg[n][0] += 1
        agg[n][1] += v
        tot += v
        seq.append((len(seq), v, n, row["Grid Size"]))
    print(f"| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| `{n}` | {c} | {v:.2f} | {100 * v / tot:.1f}% |")
    print(f"| total | {len(seq)} | {tot:.2f} | |")
    if thr is not None:
        print()
        for i, v, n, g in seq:
            if v > thr:
                print(f"{i:5d} {v:8.2f} ms  {n[:48]} {g}")


if __name__ == "__main__":
    main(sys.argv[1], thr=float(sys.argv[2]) if len(sys.argv) > 2 else None)
