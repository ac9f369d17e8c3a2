"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> share table for profiles/.
    python scripts/summarize_launches.py launches.csv "command that was profiled" > profiles/rNN_xxx_summary.txt"""
import csv
import sys
from collections import defaultdict


def main(path, cmd=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    mu = h.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        t = float(r[mv].replace(",", ""))
        t_us = t / 1e3 if r[mu] in ("ns", "nsecond") else (t if r[mu].startswith("u") else t * 1e3)
        name = r[kn].split("(")[0][:100]
        agg[name][0] += 1
        agg[name][1] += t_us
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}")
    print(f"# {n} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised: compare SHARES)")
    print("share%  launches  avg_us  kernel")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / tot:6.2f}  {c:6d}  {t / c:9.1f}  {name}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
