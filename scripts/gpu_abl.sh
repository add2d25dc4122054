#!/bin/bash
# forward-kernel ablations: the same probe with several builds of the library (AWB_LIB)
for v in "$@"; do
  echo "== $v"
  AWB_LIB=scripts/abl/lib_$v.so timeout 600 python scripts/perf_probe.py --k 50 --sites 100000 --chains 148 --reps 2 2>&1 | tail -1 | sed 's/.*| setup/setup/'
  AWB_LIB=scripts/abl/lib_$v.so timeout 600 python scripts/perf_probe.py --k 20 --sites 100000 --chains 148 --reps 2 2>&1 | tail -1 | sed 's/.*| setup/setup/'
done
