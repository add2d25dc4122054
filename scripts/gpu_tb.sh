#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py tests/test_recomb.py -m gpu -x -q 2>&1 | tail -3
python scripts/perf_probe.py --k 50 --sites 50000 --chains 148 --reps 2 --packed 1 2>&1 | grep "k=" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/' | tail -1
python scripts/perf_probe.py --k 20 --sites 50000 --chains 148 --reps 2 2>&1 | grep "k=" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/' | tail -1
