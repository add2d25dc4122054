#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_$TAG.log 2>&1
tail -9 gpurun_out/pytest_$TAG.log
for c in 3 2 1; do
  timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/bench_${TAG}_config$c.json"))
    print("config $c value %.4e e2e %.4e ms %.1f" % (l["value"], l["e2e"]["value"], l["ms_per_step"]), l["stage_ms"], l["run"]["forward_kernel"], "frac %.4f" % l["roofline"]["frac"], "fwd_only %.3e" % l["forward_only"]["value"], "h2d %.2f GB" % (l["e2e"]["h2d_bytes_per_step"]/1e9))
    print({k: round(v, 1) for k, v in l["kernel_ms"].items()}, l["roofline"]["launches_per_step"])
except Exception as e:
    print("config $c failed", e)
PY
  tail -3 gpurun_out/bench_${TAG}_config$c.err
done
