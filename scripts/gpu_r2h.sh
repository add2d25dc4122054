#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
probe() { python scripts/perf_probe.py "$@" 2>&1 | grep "k=\|forward kernel" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/' ; }
echo "== k=50 default"; AWB_VERBOSE=1 probe --k 50 --sites 50000 --chains 148 --reps 2
AWB_VERBOSE=1 probe --k 50 --sites 50000 --chains 444 --reps 2
echo "== k=50 U=2 bookw"; AWB_VERBOSE=1 AWB_K4_U=2 probe --k 50 --sites 50000 --chains 148 --reps 2
echo "== k=20 default"; AWB_VERBOSE=1 probe --k 20 --sites 50000 --chains 148 --reps 2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -x -q --durations=3 > gpurun_out/pytest_$TAG.log 2>&1
tail -4 gpurun_out/pytest_$TAG.log
