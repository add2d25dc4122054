#!/bin/bash
echo "== old tree"
(cd scripts/abl/oldrepo && timeout 600 python scripts/perf_probe.py --k 50 --sites 100000 --chains 32,148 --reps 2 2>&1 | grep "k=50" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/')
(cd scripts/abl/oldrepo && timeout 600 python scripts/perf_probe.py --k 20 --sites 100000 --chains 148 --reps 2 2>&1 | grep "k=20" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/')
echo "== current"
timeout 600 python scripts/perf_probe.py --k 50 --sites 100000 --chains 32,148 --reps 2 2>&1 | grep "k=50" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/'
timeout 600 python scripts/perf_probe.py --k 20 --sites 100000 --chains 148 --reps 2 2>&1 | grep "k=20" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/'
