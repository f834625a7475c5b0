#!/bin/bash
# K7 (ts) ablations on a -DPMD_TUNE build: 1 no MMAs, 2 no drains, 4 no conversion, 8 no movie loads (results wrong)
for a in ${ABL:-0 1 2 3 4 8 10 12 14 15}; do
  echo "== ABLATE $a"
  PMD_TS_ABLATE=$a python scripts/bench_k7.py 20000 3 ts 2>&1 | grep -E "project_stream_ts"
done
