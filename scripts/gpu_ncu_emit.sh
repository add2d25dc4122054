#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'awb_emit' -c 1 -o /tmp/emit -f \
    python scripts/perf_probe.py --k 50 --sites 20000 --chains 148 --reps 1 --checkpoint 1 --packed 1 > gpurun_out/ncu_emit.log 2>&1
ncu -i /tmp/emit.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_emit.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_emit.csv > gpurun_out/hot_lines_awb_emit.txt 2>&1
ncu -i /tmp/emit.ncu-rep --page raw --csv > /tmp/emit_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/emit_raw.csv > gpurun_out/ncu_emit_summary.txt 2>&1
cat gpurun_out/ncu_emit_summary.txt | head -30
awk 'NR==1 || NR%2==0' gpurun_out/hot_lines_awb_emit.txt | head -40
tail -3 gpurun_out/ncu_emit.log
