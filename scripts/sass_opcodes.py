"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libpmd_sm100.so (tcgen05 MMAs = UTC*MMA, tensor-memory
loads / stores = LDTM / STTM, TMA = UTMALDG / UBLKCP, packed fp32 = FADD2 / FMUL2, FP64 tensor = DMMA).
Usage: python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "localmd_b200", "libpmd_sm100.so")], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "FADD2", "FMUL2", "FFMA2",
         "DMMA", "HMMA", "LDGSTS", "RED", "ATOMG"]
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                hist[kern][w] += 1
print("SASS opcode counts per kernel of libpmd_sm100.so (cuobjdump -sass; sm_100a).  Only kernels with tensor-core / TMA / "
      "tensor-memory instructions are listed in full; the rest as totals.\n")
for k, h in hist.items():
    special = {w: c for w, c in h.items() if w != "total" and w not in ("SYNCS", "RED", "ATOMG", "LDGSTS")}
    if special:
        print("%-70s %6d instr  %s" % (k[:70], h["total"], "  ".join("%s %d" % (w, h[w]) for w in WATCH if h[w])))
print()
for k, h in hist.items():
    special = {w: c for w, c in h.items() if w != "total" and w not in ("SYNCS", "RED", "ATOMG", "LDGSTS")}
    if not special:
        print("%-70s %6d instr  %s" % (k[:70], h["total"], "  ".join("%s %d" % (w, h[w]) for w in WATCH if h[w])))
