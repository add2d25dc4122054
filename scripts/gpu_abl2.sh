#!/bin/bash
# co-resident CTAs: more windows than SMs
for v in "$@"; do
  echo "== $v"
  AWB_LIB=scripts/abl/lib_$v.so timeout 600 python scripts/perf_probe.py --k 50 --sites 100000 --chains 148,296 --reps 2 2>&1 | grep "k=50" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/'
  AWB_LIB=scripts/abl/lib_$v.so timeout 600 python scripts/perf_probe.py --k 20 --sites 100000 --chains 148,296,444 --reps 2 2>&1 | grep "k=20" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/'
done
