#!/bin/bash
# tests + stage probe + per-kernel launch list on a mid-size batch
mkdir -p gpurun_out
TAG=${1:-q}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python scripts/perf_probe.py --k 50 --sites 200000 --chains 32 --reps 2 2>&1 | tail -2
timeout 600 python scripts/perf_probe.py --k 50 --sites 200000 --chains 32 --internal 1 --reps 1 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_$TAG.csv \
    python scripts/perf_probe.py --k 50 --sites 200000 --chains 32 --reps 1 > /dev/null 2>&1
python - <<PY
import csv
for r in csv.reader(open("gpurun_out/launches_$TAG.csv")):
    if len(r) > 14 and r[0].isdigit():
        print("%-70s %10.3f ms" % (r[4][:70], float(r[14].replace(",", "")) / 1e6))
PY
