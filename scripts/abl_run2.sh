#!/bin/bash
for k in "$@"; do
  echo "== $k"
  AWB_LIB=scripts/abl/lib_$k.so timeout 300 python scripts/perf_probe.py --k 50 --sites 200000 --chains 32 --reps 1 2>&1 | tail -2 | sed 's/.*| setup/setup/'
done
