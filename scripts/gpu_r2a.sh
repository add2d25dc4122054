#!/bin/bash
# round-2 first visit: whole gpu test suite (incl. at-size parity), the default bench and the other configs
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_$TAG.log 2>&1
tail -25 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for c in 1 2 4 5; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
  cat gpurun_out/bench_${TAG}_config$c.json; tail -3 gpurun_out/bench_${TAG}_config$c.err
done
