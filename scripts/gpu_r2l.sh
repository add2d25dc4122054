#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2l}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py tests/test_recomb.py tests/test_unphased.py tests/test_gpu_packed.py -m gpu -x -q --durations=3 > gpurun_out/pytest_$TAG.log 2>&1
tail -6 gpurun_out/pytest_$TAG.log
run() {
  python - "$1" <<PY
import json, sys
try:
    l = json.load(open(sys.argv[1]))
    print("value %.4e ms %.1f" % (l["value"], l["ms_per_step"]), {k: round(v) for k, v in l["stage_ms"].items()}, "frac %.4f" % l["roofline"]["frac"], {k: round(v, 1) for k, v in l["kernel_ms"].items()}, l["e2e"] and "e2e %.4e" % l["e2e"]["value"])
except Exception as e:
    print("failed", e)
PY
}
timeout 900 python bench.py --config 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_config3.json 2> gpurun_out/bench_${TAG}_config3.err; run gpurun_out/bench_${TAG}_config3.json
tail -2 gpurun_out/bench_${TAG}_config3.err
AWB_NO_OVERLAP=1 timeout 900 python bench.py --config 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_noov.json 2>/dev/null; run gpurun_out/bench_${TAG}_noov.json
timeout 600 compute-sanitizer --tool synccheck --print-limit 200 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_vectors or checkpointed_small" \
    > gpurun_out/sanitizer_synccheck_$TAG.log 2>&1
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_synccheck_$TAG.log | tail -3
grep "at \|Device Frame" gpurun_out/sanitizer_synccheck_$TAG.log | sort | uniq -c | sort -rn | head -8
grep "by thread" gpurun_out/sanitizer_synccheck_$TAG.log | awk '{print $4}' | sort | uniq -c | sort -rn | head -40 | tr '\n' ' '
