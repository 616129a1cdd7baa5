"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the
LAST decode step (from the last embed_kernel launch to the end).  usage:
  python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("acp::", "").replace("<unnamed>::", "").replace("void ", "") for r in rows]
    durs = [float(r["Metric Value"]) / 1000.0 for r in rows]
    starts = [i for i, n in enumerate(names) if "embed_kernel" in n]
    start = starts[-1]
    agg = collections.OrderedDict()
    for n, d, r in zip(names[start:], durs[start:], rows[start:]):
        a = agg.setdefault(f"{n} grid={r['Grid Size']} block={r['Block Size']}", [0, 0.0])
        a[0] += 1
        a[1] += d
    tot = sum(durs[start:])
    print(f"source: {path}  ({len(rows)} launches captured; engine steps start at launches {starts})\n")
    print(f"last decode step: {len(names) - start} launches, sum of gpu__time_duration = {tot:.1f} us "
          "(cold-cache, serialised under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.2f} |")


if __name__ == "__main__":
    main(sys.argv[1])
