#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_infsites.py tests/test_unphased.py tests/test_gpu_packed.py tests/test_gpu_compat.py tests/test_gpu_dropin.py -m gpu -x -q --durations=3 > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --config 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_config3.json 2> gpurun_out/bench_${TAG}_config3.err
python - <<PY
import json
l = json.load(open("gpurun_out/bench_${TAG}_config3.json"))
print("value %.4e ms %.1f" % (l["value"], l["ms_per_step"]), {k: round(v) for k, v in l["stage_ms"].items()}, "frac %.4f" % l["roofline"]["frac"], {k: round(v, 1) for k, v in l["kernel_ms"].items()})
PY
