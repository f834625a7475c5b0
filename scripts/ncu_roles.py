"""Stall-sample breakdown of an ncu report's source page: top sampled SASS lines with their dominant stall reason.
Usage: python scripts/ncu_roles.py report.ncu-rep [top_n] [lo hi]"""
import csv
import subprocess
import sys


def main(path, top=40, lo=0, hi=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    hi = len(data) if hi is None else hi
    print("total samples", sum(int(r[iS]) for r in data), "warp instructions", sum(int(r[iEx]) for r in data))
    idx = sorted(range(lo, hi), key=lambda i: -int(data[i][iS]))[:top]
    for i in sorted(idx):
        r = data[i]
        st = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:1]
        print(i, r[iS], r[iEx], r[iSrc].strip()[:90], st)


if __name__ == "__main__":
    a = sys.argv
    main(a[1], int(a[2]) if len(a) > 2 else 40, int(a[3]) if len(a) > 3 else 0, int(a[4]) if len(a) > 4 else None)
