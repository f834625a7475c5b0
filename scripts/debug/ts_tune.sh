#!/bin/bash
# K7 (ts) access-shape experiments on a -DPMD_TUNE build: L2 promotion of the TMA boxes x cache policy of the movie
for promo in 256 128 0; do for pol in 0 1; do
  echo "== PROMO $promo POLICY $pol"
  PMD_TS_PROMO=$promo PMD_TS_POLICY=$pol python scripts/bench_k7.py 20000 3 ts 2>&1 | grep -E "project_stream_ts|vs float64"
done; done
