#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_$TAG.log 2>&1
tail -9 gpurun_out/pytest_$TAG.log
for c in 3; do
  AWB_VERBOSE=1 timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/bench_${TAG}_config$c.json"))
    print("config $c value %.4e e2e %.4e ms %.1f" % (l["value"], l["e2e"]["value"], l["ms_per_step"]), l["stage_ms"], l["run"]["forward_kernel"], "frac %.4f" % l["roofline"]["frac"], "fwd_only %.3e" % l["forward_only"]["value"])
    print(l["run"]["table"])
except Exception as e:
    print("config $c failed", e)
PY
  grep "forward kernel\|segment tables" gpurun_out/bench_${TAG}_config$c.err | sort | uniq -c | head
  tail -3 gpurun_out/bench_${TAG}_config$c.err
done
