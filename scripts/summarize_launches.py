"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1 :]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"nsecond": v / 1e6, "ns": v / 1e6, "usecond": v / 1e3, "us": v / 1e3}.get(r[ui], v)
        a = agg.setdefault(r[ki][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("launches %d, summed device time %.1f ms (ncu: cold-cache, serialised -- compare shares)" % (len(data), tot))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%9.2f ms %5.1f%%  x%-4d %s" % (v, 100 * v / tot, c, n))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
