#!/bin/bash
TAG=${1:-x}
KREGEX=${2:-awb_forward_fast}
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:$KREGEX -c 1 -o gpurun_out/ncu_$TAG -f \
    python scripts/perf_probe.py --k 50 --sites 100000 --chains 32 --reps 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
