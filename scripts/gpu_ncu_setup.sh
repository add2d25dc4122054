#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'awb_block_setup|awb_switch_setup|awb_emit' -c 3 -o /tmp/setup -f \
    python scripts/perf_probe.py --k 50 --sites 20000 --chains 148 --reps 1 > gpurun_out/ncu_setup.log 2>&1
for k in awb_block_setup awb_switch_setup awb_emit; do
  ncu -i /tmp/setup.ncu-rep --page source --print-source cuda,sass --csv -k regex:$k > /tmp/src_$k.csv 2>/dev/null
  python scripts/ncu_lines.py /tmp/src_$k.csv > gpurun_out/hot_lines_$k.txt 2>&1
done
ncu -i /tmp/setup.ncu-rep --page raw --csv > /tmp/setup_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/setup_raw.csv > gpurun_out/ncu_setup_summary.txt 2>&1
head -60 gpurun_out/hot_lines_awb_block_setup.txt
