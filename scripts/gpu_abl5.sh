#!/bin/bash
for v in "$@"; do
  echo "== $v"
  AWB_LIB=scripts/abl/lib_$v.so timeout 600 python scripts/perf_probe.py --k 50 --sites 100000 --chains 32 --reps 2 2>&1 | grep "k=50" | tail -1 | sed 's/.*C=/C=/; s/| gen.*| setup/setup/'
done
