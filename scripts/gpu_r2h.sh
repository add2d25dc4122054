#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -x -q --durations=3 > gpurun_out/pytest_$TAG.log 2>&1
tail -8 gpurun_out/pytest_$TAG.log
probe() { python scripts/perf_probe.py "$@" 2>&1 | grep "k=" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/' | awk 'NR%2==0'; }
echo "== k=50 U=4 (96 regs)"; AWB_K4_U=4 probe --k 50 --sites 50000 --chains 148,296,444 --reps 2
echo "== k=50 U=2"; AWB_K4_U=2 probe --k 50 --sites 50000 --chains 148,296 --reps 2
echo "== k=50 U=1"; AWB_K4_U=1 probe --k 50 --sites 50000 --chains 148 --reps 2
echo "== k=20 U=2"; AWB_K4_U=2 probe --k 20 --sites 50000 --chains 148,592 --reps 2
echo "== k=20 U=1"; AWB_K4_U=1 probe --k 20 --sites 50000 --chains 148,296 --reps 2
