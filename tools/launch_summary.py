"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST device-resident
timed step of bench.py (delimited by the 256 MB L2-flush fills).  Usage: python tools/launch_summary.py launches.csv [top]"""
import collections
import csv
import re
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [r[kn] for r in data]
    t = [float(r[mv].replace(",", "")) for r in data]
    fills = [i for i, n in enumerate(names) if "FillFunctor<unsigned char>" in n]
    a, b = fills[-2], fills[-1]           # [.. warm-up .., flush, TIMED STEP, flush, e2e step]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for i in range(a + 1, b):
        n = re.sub(r"\(.*", "", names[i])
        n = re.sub(r"^void ", "", n)[:80]
        agg[n][0] += 1
        agg[n][1] += t[i]
    tot = sum(v[1] for v in agg.values())
    print(f"timed step: {b - a - 1} launches, {tot / 1e6:.2f} ms summed kernel time (ncu: serialised, cold cache)")
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{s / 1e6:8.3f} ms {100 * s / tot:5.1f}%  x{c:<3d} {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
