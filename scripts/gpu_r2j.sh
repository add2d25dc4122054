#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_$TAG.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[hdr+1:]:
    if len(r) <= vi: continue
    k = r[ki].split("(")[0][:60]
    tot[k] += float(r[vi].replace(",", "")); cnt[k] += 1
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-62s n=%4d  %9.2f ms  %5.1f %%" % (k, cnt[k], v / 1e6, 100 * v / T))
PY
