#!/bin/bash
for u in 1 2 4; do
AWB_VERBOSE=1 AWB_K4_U=$u timeout 300 python scripts/perf_probe.py --k 50 --sites 5000 --chains 2 --reps 1 2>&1 | grep "forward kernel"
AWB_VERBOSE=1 AWB_K4_U=$u timeout 300 python scripts/perf_probe.py --k 20 --sites 5000 --chains 2 --reps 1 2>&1 | grep "forward kernel"
done
AWB_VERBOSE=1 timeout 300 python scripts/perf_probe.py --k 100 --ntimes 40 --sites 5000 --chains 2 --reps 1 2>&1 | grep "forward kernel"
