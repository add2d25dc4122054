#!/bin/bash
# One GPU-box visit: tests, bench (both arms), ncu launch list, dram traffic of the
# forward kernel at bench size, one --set full capture at a reduced size.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:awb_forward_fast -c 60 --csv \
    --log-file gpurun_out/traffic_fwd_$TAG.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traffic_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'awb_forward_fast|awb_traceback|awb_emit' -c 6 \
    -o gpurun_out/full_$TAG -f \
    python bench.py --steps 1 --warmup 0 --windows 8 --sites 100000 --no-cpu-baseline --no-e2e \
    > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -12
