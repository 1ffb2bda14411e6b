"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def main(path, top=40):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"void |hg::|at::native::|<unnamed>::", "", name)[:64]
        t = float(r[vi].replace(",", ""))
        t = t / 1000 if r[ui] == "ns" else (t * 1000 if r[ui] == "ms" else t)
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1000:.3f} ms of kernel time (serialised, cold cache)")
    print(f"# {'kernel':64s} {'calls':>6s} {'total ms':>9s} {'avg us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:66s} {v[0]:6d} {v[1] / 1000:9.3f} {v[1] / v[0]:8.2f} {v[1] / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
