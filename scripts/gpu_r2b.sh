#!/bin/bash
# drop-in tests, the small-population test, config-1/2 bench lines with the MCMC pair
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q --durations=8 > gpurun_out/pytest_$TAG.log 2>&1
tail -30 gpurun_out/pytest_$TAG.log
for c in 1 2; do
  timeout 900 python bench.py --config $c > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
  python - <<PY
import json
l = json.load(open("gpurun_out/bench_${TAG}_config$c.json"))
print("config $c value %.3e e2e %.3e" % (l["value"], l["e2e"]["value"]), json.dumps(l.get("mcmc")))
PY
  tail -3 gpurun_out/bench_${TAG}_config$c.err
done
