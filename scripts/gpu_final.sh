#!/bin/bash
# last GPU-box visit of the round: the whole GPU suite, the smoke entry, per-kernel
# times of the probe shape, the headline bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1
tail -3 gpurun_out/pytest_final.log
timeout 200 python scripts/perf_probe.py --k 50 --sites 50000 --chains 148 --reps 2 --packed 1 --ktimes 1 2>&1 | grep "k=\|kernel ms" | sed 's/.*C=/C=/; s/| gen.*| setup/setup/' | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e ms %.1f frac %.4f" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["roofline"]["frac"]), l.get("kernel_ms"), l.get("parity"), l.get("clocks"))
PY
