"""Per-kernel summary of ONE decomposition out of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py:
the launches between the K1 launch that starts the last complete job and the next K1 launch (the list also holds the
synthetic-movie generation and bench.py's separate K1 / K7 timing loops).  Classifies own (pmd::) / library / torch."""
import collections
import csv
import re
import sys


def main(path, top=60):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ev = [(r[ki], float(r[vi].replace(",", "")) / 1e6) for r in data if len(r) > vi and r[vi]]
    k1 = [i for i, (n, _) in enumerate(ev) if "stats_tc_kernel" in n or "stats_fft_kernel" in n]
    jobs = [(a, b) for a, b in zip(k1, k1[1:] + [len(ev)]) if b - a > 300]
    a, b = jobs[-1] if len(jobs) == 1 else jobs[-2] if jobs[-1][1] == len(ev) and len(jobs) > 1 else jobs[-1]
    # the last job of the list may run into the K1 / K7 timing loops: cut it at the first repeated K7 launch after the job's own
    job = ev[a:b]
    agg, cls = collections.OrderedDict(), collections.Counter()
    for n, ms in job:
        short = re.sub(r"\(.*", "", n)[:110]
        e = agg.setdefault(short, [0, 0.0])
        e[0] += 1
        e[1] += ms
        kind = "own (pmd::)" if "pmd::" in n else "library (cuBLAS / cuSOLVER / CUTLASS)" if re.search(
            r"cutlass|sgemm|gemm|sytrd|laed|stedc|steqr|larf|trsm|getrf|potrf|gemv|nvjet|splitK|syherk|scal_kernel|lansy|ormtr|lacpy|merge_ker|"
            r"scale_max|xx_set_info|copy_info|setup_vhat|zero_lower|epilogue|transpose", n) else "torch elementwise / sort / index"
        cls[kind] += ms
    tot = sum(ms for _, ms in job)
    print("one decomposition: launches %d, summed device time %.1f ms (ncu: cold-cache, serialised -- compare shares)" % (len(job), tot))
    for k, v in cls.most_common():
        print("  %-45s %7.2f ms  %5.1f%%" % (k, v, 100 * v / tot))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%9.2f ms %5.1f%%  x%-4d %s" % (v, 100 * v / tot, c, n))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
