#!/bin/bash
for k in "$@"; do
  echo "== ablate $k"
  AWB_LIB=scripts/abl/lib_$k.so timeout 300 python scripts/perf_probe.py --k 50 --sites 100000 --chains 32 --reps 2 2>&1 | tail -1 | sed 's/.*| setup/setup/'
done
