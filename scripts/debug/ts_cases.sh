#!/bin/bash
# every parametrisation of test_project_stream_ts in its own process (a CUDA fault poisons the context)
ids=$(python -m pytest tests/test_kernels_gpu.py --collect-only -q -k project_stream_ts 2>/dev/null | grep "::")
for id in $ids; do
  timeout 180 python -m pytest "$id" -x -q > /tmp/one.log 2>&1
  rc=$?
  echo "$rc $id $(grep -E 'illegal|misaligned|stuck|Mismatch|AssertionError|Max abs|passed|failed' /tmp/one.log | head -3 | tr '\n' ' ')"
done
