#!/bin/bash
# K1 (block set-up) with its node arrays staged in shared memory: parity, then kernel times
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -x -q 2>&1 | tail -3
for ip in 0 1; do
  if [ $ip = 1 ]; then export AWB_K1_IN_PLACE=1; else unset AWB_K1_IN_PLACE; fi
  echo "== in_place=$ip"
  timeout 200 python scripts/perf_probe.py --k 50 --sites 50000 --chains 148 --reps 2 --packed 1 --ktimes 1 2>&1 | grep "k=\|kernel ms" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/' | tail -2
done
